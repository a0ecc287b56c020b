#!/bin/bash
# Round-2 addendum: `ncu --set full` captures of the kernels added late in the round (fused GEMM + LayerNorm epilogues, LAS stacked
# key/value gradients) from one eager step; summaries appended to gpurun_out/r02b_ncu_summary.txt
set -u
O=gpurun_out
mkdir -p $O
cap() { ncu --set full --clock-control none -k regex:$1 -s $2 -c 1 -f -o $O/r02b_$1 python scripts/one_step.py 2 bf16 > /dev/null 2>&1; }
cap gemm_ln_kernel 40
cap gemm_lnbwd_kernel 40
cap las_stack_grad 2
for f in $O/r02b_*.ncu-rep; do
  b=$(basename $f .ncu-rep)
  ncu -i $f --page raw --csv > $O/$b.csv 2>/dev/null
  echo "== $b"; python scripts/ncu_summary.py $O/$b.csv
done > $O/r02b_ncu_summary.txt 2>&1
rm -f $O/r02b_*.csv $O/r02b_*.ncu-rep
cat $O/r02b_ncu_summary.txt | grep -E "^==|time_duration|dram__bytes|utchmma|cluster_size|registers"
