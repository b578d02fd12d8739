"""Turn the ncu artefacts of profiles/collect.sh into the committed text summaries.
usage: python profiles/summarize.py r01   (reads gpurun_out/r01_*, writes profiles/r01_*.md / .csv)"""
import csv
import os
import shutil
import subprocess
import sys

R = sys.argv[1] if len(sys.argv) > 1 else 'r01'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO, PR = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed']


def launches(name):
    src = os.path.join(GO, f'{R}_launches_{name}.csv')
    if not os.path.exists(src):
        return
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ik, iv = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = {}
    for r in rows[1:]:
        k = r[ik].split('(')[0][:70]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv].replace(',', ''))
    tot = sum(v[1] for v in agg.values())
    unit = rows[1][hdr.index('Metric Unit')]
    with open(os.path.join(PR, f'{R}_launches_{name}.md'), 'w') as f:
        f.write(f'# {R}: launches of `python bench.py` ({name}) under `ncu --metrics gpu__time_duration.sum`\n\n')
        f.write('Cold-cache, serialised timings: compare shares, not absolutes.\n\n')
        if name in ('wave', 'bench'):
            f.write('NOTE: in the real step the boundary-row launch on the side stream (`jet_simt_kernel` on a few CTAs: periodic / '
                    'finite-difference groups) runs CONCURRENTLY with the interior launch and ends before it (DESIGN 3.4); ncu '
                    'serialises them, so its share below is not a share of the step.  The bench command measures every '
                    'BASELINE config one after the other: kernels are listed over the whole run.\n\n')
        f.write('| kernel | launches | total | share |\n|---|---|---|---|\n')
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'| `{k}` | {n} | {t:.1f} {unit} | {100 * t / tot:.1f} % |\n')
    shutil.copy(src, os.path.join(PR, f'{R}_launches_{name}.csv'))


def full(name):
    rep = os.path.join(GO, f'{R}_{name}.ncu-rep')
    if not os.path.exists(rep):
        return
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(os.path.join(PR, f'{R}_ncu_{name}.md'), 'w') as f:
        f.write(f'# {R}: `ncu --set full --clock-control none` of `{vals[hdr.index("Kernel Name")][:80]}`\n\n| metric | value | unit |\n|---|---|---|\n')
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                f.write(f'| {k} | {vals[i]} | {units[i]} |\n')
        f.write('\n## warp stall reasons (per issue-active)\n\n')
        for i, h in enumerate(hdr):
            if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio'):
                try:
                    if float(vals[i]) > 0.05:
                        f.write(f'* {h.split("stalled_")[1].split("_per_issue")[0]}: {float(vals[i]):.2f}\n')
                except ValueError:
                    pass
        src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source=cuda,sass'],
                             capture_output=True, text=True).stdout
        tmp = os.path.join(GO, f'{R}_{name}_src.csv')
        open(tmp, 'w').write(src)
        hot = subprocess.run([sys.executable, os.path.join(PR, 'hotlines.py'), tmp, '20'], capture_output=True, text=True).stdout
        f.write('\n## hottest source lines (warp-stall samples)\n\n```\n' + hot + '```\n')


import json
for n in ('wave', 'bench', 'mat'):
    launches(n)
for n in ('jet_tc', 'jet_simt', 'jet_tcs_cfg1', 'jet_tcs', 'jet_tcs_ns', 'wgrad', 'mat'):
    full(n)

# dram traffic per launch of the dominant kernel of each bench workload, read by bench.py (roofline.traffic)
traffic = {}
for workload, reps in (('burgers_NN_cfg1', ['jet_tcs_cfg1']), ('wave_autograd_1e6', ['jet_tcs', 'wgrad']),
                       ('ns_autograd_1e6', ['jet_tcs_ns']), ('poisson_mat_4096', ['mat'])):
    tot, names = 0.0, []
    for name in reps:
        rep = os.path.join(GO, f'{R}_{name}.ncu-rep')
        if not os.path.exists(rep):
            tot = None
            break
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units, vals = rows[0], rows[1], rows[2]
        scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
        for key in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            i = hdr.index(key)
            tot += float(vals[i].replace(',', '')) * scale[units[i]]
        names.append(vals[hdr.index('Kernel Name')].split('(')[0])
    if tot is not None:
        traffic[workload] = {'bytes': tot, 'kernels': names, 'source': f'profiles/{R}_ncu_*.md (ncu --set full, one launch each)'}
if traffic:
    path = os.path.join(PR, 'ncu_traffic.json')
    old = json.load(open(path)) if os.path.exists(path) else {}
    old.update(traffic)
    json.dump(old, open(path, 'w'), indent=1)
print('\n'.join(sorted(x for x in os.listdir(PR) if x.startswith(R))))
