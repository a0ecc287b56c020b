"""Pull the handful of metrics DESIGN.md / bench.py quote out of an `ncu --page raw --csv` dump.
usage: ncu -i X.ncu-rep --page raw --csv > X.csv ; python scripts/ncu_summary.py X.csv"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
h, u, v = rows[0], rows[1], rows[2]
KEYS = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_uniform', 'launch__grid_size', 'launch__block_size', 'launch__cluster_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.max',
        'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'lts__t_bytes.sum', 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed']
out = {}
for i, name in enumerate(h):
    short = name.split('TriageCompute.')[-1]
    if short in KEYS:
        out[short] = (v[i], u[i])
for k in KEYS:
    if k in out:
        print(f'{k:100s} {out[k][0]} {out[k][1]}')
