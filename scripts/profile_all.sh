#!/bin/bash
# Round evidence for profiles/: launch list of one eager step, full ncu captures of the dominant kernels, the in-kernel
# recurrence timeline, per-section graph timings.  Run on the GPU box:  bash scripts/profile_all.sh
set -u
O=gpurun_out
mkdir -p $O
timeout 200 python scripts/one_step.py 1 bf16 > $O/one_step.log 2>&1 || { echo "one_step failed"; tail -5 $O/one_step.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r1.csv python scripts/one_step.py 1 bf16 > $O/ncu_list.log 2>&1
timeout 100 python scripts/profile_blstm.py 1008 64 > $O/blstm_timeline.txt 2>&1
for k in blstm_fwd_tc blstm_bwd_tc; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o $O/r01_$k python scripts/profile_blstm.py 1008 64 > /dev/null 2>&1
done
timeout 100 python scripts/profile_gemm.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o $O/r01_gemm_tc python scripts/profile_gemm.py > /dev/null 2>&1
for k in mha_fwd_tc mha_bwd_tc las_attn_fwd_cl gemm_tc_clk; do
  ncu --set full --clock-control none -k regex:$k -s 40 -c 1 -f -o $O/r01_$k python scripts/one_step.py 1 bf16 > /dev/null 2>&1
done
timeout 300 python scripts/section_times.py > $O/section_times.txt 2>&1
ls -la $O/*.ncu-rep; tail -6 $O/section_times.txt
