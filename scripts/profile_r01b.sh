#!/bin/bash
# Second evidence pass of round 1 (after the persistent / CTA-pair GEMM, fused optimizer, dropout and cached decode):
# launch list of one eager training step incl. clip + Adam, and `ncu --set full` captures of the tensor-bound GEMM and of
# the HBM-bound kernels the north star names (loss, LayerNorm, mix) plus the optimizer.  Run on the GPU box.
set -u
O=gpurun_out
mkdir -p $O
timeout 200 python scripts/one_step.py 2 bf16 > $O/one_step_b.log 2>&1 || { echo "one_step failed"; tail -5 $O/one_step_b.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_r1b.csv python scripts/one_step.py 2 bf16 > $O/ncu_list_b.log 2>&1
python scripts/summarize_launches.py $O/launches_r1b.csv > $O/launches_r1b_summary.txt 2>&1
timeout 100 python scripts/profile_gemm.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_tc_pair -s 2 -c 1 -f -o $O/r01b_gemm_tc_pair python scripts/profile_gemm.py > /dev/null 2>&1
cap() {  # kernel regex, launches to skip
  ncu --set full --clock-control none -k regex:$1 -s $2 -c 1 -f -o $O/r01b_$1 python scripts/one_step.py 2 bf16 > /dev/null 2>&1
}
cap softmax_nll_fused 2
cap layernorm_fwd 70
cap layernorm_bwd_reg 70
cap mix_gather_concat 2
cap multi_adam 2
cap multi_sqnorm 2
cap mha_fwd_tc 40
cap mha_bwd_tc 40
for f in $O/r01b_*.ncu-rep; do
  b=$(basename $f .ncu-rep)
  ncu -i $f --page raw --csv > $O/$b.csv 2>/dev/null
  echo "== $b"; python scripts/ncu_summary.py $O/$b.csv
done > $O/r01b_ncu_summary.txt 2>&1
rm -f $O/r01b_*.csv
ls -la $O/r01b_*.ncu-rep | wc -l; head -12 $O/launches_r1b_summary.txt
