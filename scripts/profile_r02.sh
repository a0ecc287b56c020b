#!/bin/bash
# Round-2 evidence pass (run on the GPU box, after the plain bench has exited 0): ncu launch list of ONE REPLAY OF THE WHOLE-STEP
# CUDA GRAPH (the thing bench.py times), and `ncu --set full` captures of the dominant kernel pair (BLSTM recurrence, for
# roofline.traffic), the persistent LAS decoder forward, the CTA-pair GEMM, the tcgen05 attention kernels and the HBM-bound
# kernels the north star names.  Summaries -> gpurun_out/r02_*.txt (copied to profiles/ by hand).
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python scripts/one_step_graph.py > $O/one_step_graph.log 2>&1 || { echo "one_step_graph failed"; tail -5 $O/one_step_graph.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_graph_launches.csv python scripts/one_step_graph.py > $O/ncu_list_graph.log 2>&1
echo "graph launch list: $(grep -c '^"' $O/r02_graph_launches.csv) rows"
for k in blstm_fwd_tc blstm_bwd_tc; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o $O/r02_$k python scripts/profile_blstm.py 1008 64 > /dev/null 2>&1
done
timeout 100 python scripts/profile_gemm.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_tc_pair -s 2 -c 1 -f -o $O/r02_gemm_tc_pair python scripts/profile_gemm.py > /dev/null 2>&1
cap() {  # kernel regex, launches to skip
  ncu --set full --clock-control none -k regex:$1 -s $2 -c 1 -f -o $O/r02_$1 python scripts/one_step.py 2 bf16 > /dev/null 2>&1
}
cap las_dec_fwd_persist 2
cap softmax_nll_fused 2
cap layernorm_fwd_vec 70
cap layernorm_bwd_reg 70
cap mix_gather_concat 2
cap multi_adam 2
cap mha_fwd_tc 40
cap mha_bwd_tc 40
for f in $O/r02_*.ncu-rep; do
  b=$(basename $f .ncu-rep)
  ncu -i $f --page raw --csv > $O/$b.csv 2>/dev/null
  echo "== $b"; python scripts/ncu_summary.py $O/$b.csv
done > $O/r02_ncu_summary.txt 2>&1
rm -f $O/r02_blstm*.csv $O/r02_gemm*.csv $O/r02_las*.csv $O/r02_soft*.csv $O/r02_layer*.csv $O/r02_mix*.csv $O/r02_multi*.csv $O/r02_mha*.csv
ls $O/r02_*.ncu-rep | wc -l; grep -c "==" $O/r02_ncu_summary.txt
