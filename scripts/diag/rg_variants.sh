mkdir -p gpurun_out
for v in 0 1 2 4 5 7; do
  echo "=== B200ST_RG_VAR=$v"
  B200ST_RG_VAR=$v timeout 200 python scripts/profile_blstm.py 1008 64 2>&1 | grep -A3 "backend auto" | tail -2
  B200ST_RG_VAR=$v timeout 200 python scripts/profile_blstm.py 1008 64 2>&1 | grep -E "^fwd_rg cycles|^bwd_rg cycles"
done > gpurun_out/s3_rg_variants.txt 2>&1
cat gpurun_out/s3_rg_variants.txt
