# N = 2 sweep of NCCL settings (results: profiles/r02_dp_timeline_n2.txt).  # N = 2 sweep of NCCL settings for the data-parallel step (bench.py under torchrun): which knob moves the +0.4 ms?
mkdir -p gpurun_out
run() {  # label, env...
  label=$1; shift
  env B200ST_DP_FLAT=0 "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/n2_$label.json 2> gpurun_out/n2_$label.err
  python -c "import json;d=json.load(open('gpurun_out/n2_$label.json'));print('$label', round(d['ms_per_step'],3), round(d['fwd_bwd_only']['ms_per_step'],3))" || tail -3 gpurun_out/n2_$label.err
}
run default
run minch16 NCCL_MIN_NCHANNELS=16
run minch32 NCCL_MIN_NCHANNELS=32
run ll128 NCCL_PROTO=LL128
run simple NCCL_PROTO=Simple
run tree NCCL_ALGO=Tree
run nthreads256 NCCL_NTHREADS=256
run buf8m NCCL_BUFFSIZE=8388608
