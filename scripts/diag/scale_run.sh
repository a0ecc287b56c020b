# bench.py at N GPUs of one box (driver-style launch); usage: bash scripts/diag/scale_run.sh N
N=$1
mkdir -p gpurun_out
if [ "$N" = 1 ]; then
  timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench_n1.err
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02e_bench_n$N.json 2> gpurun_out/r02e_bench_n$N.err
fi
python -c "
import json
for l in open('gpurun_out/r02e_bench_n$N.json'):
    if l.startswith('{'):
        d=json.loads(l);print('N=$N', round(d['ms_per_step'],3), round(d['value'],1), d['clocks'])" || tail -5 gpurun_out/r02e_bench_n$N.err
