"""Aggregate an `ncu --page source --csv --print-source=cuda,sass` export by CUDA source line.
usage: ncu -i X.ncu-rep --page source --csv --print-source=cuda,sass > x.csv; python profiles/hotlines.py x.csv [N]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur_file = None
agg = collections.OrderedDict()
tot = totin = 0
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if len(r) < 8 or r[0] in ('Line No', ''):
        continue
    try:
        ln = int(r[0]); s = int(r[6] or 0); ins = int(r[7] or 0)
    except ValueError:
        continue
    a = agg.setdefault((cur_file, ln), [0, 0, r[1].strip()[:100]])
    a[0] += s; a[1] += ins
    tot += s; totin += ins
print('total samples', tot, 'warp instructions', totin)
for (f, ln), (s, ins, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f'{f}:{ln:4d} {100 * s / tot:5.1f}% smp {100 * ins / totin:5.1f}% ins | {src}')
