"""Print the headline counters, stall reasons and hottest source lines of one .ncu-rep.
usage: python profiles/ncu_summary.py gpurun_out/X.ncu-rep [n_lines]"""
import csv, io, subprocess, sys, os
rep = sys.argv[1]; top = sys.argv[2] if len(sys.argv) > 2 else '30'
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[-1]
d = dict(zip(hdr, vals))
keys = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_bytes.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__inst_executed_pipe_lsu.sum', 'smsp__inst_executed_op_shared_st.sum', 'smsp__inst_executed_op_global_st.sum']
for k in keys:
    if k in d: print(f'{k:90s} {d[k]} {rows[1][hdr.index(k)] if len(rows) > 2 else ""}')
print('stall reasons (per issue):')
for h in hdr:
    if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio'):
        try:
            v = float(d[h])
        except ValueError:
            continue
        if v > 0.15: print('   ', h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')], round(v, 2))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source=cuda,sass'], capture_output=True, text=True).stdout
open('/tmp/_src.csv', 'w').write(src)
subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), 'hotlines.py'), '/tmp/_src.csv', top])
