"""Instructions executed per CUDA source line of an ncu report (needs -lineinfo and --import-source on).
usage: python scripts/ncu_lines.py report.ncu-rep [elements_per_launch] [top_n]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
nelem = float(sys.argv[2]) if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr = '?', None
agg = collections.defaultdict(lambda: [0, 0, ''])
for r in rows:
    if r and r[0] in ('File Name', 'File Path'):
        cur, hdr = r[1], None
        continue
    if r and r[0] == 'Line No':
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():          # per-line totals (SASS rows have an empty line number)
        try:
            inst, samp = int(r[hdr.index('Instructions Executed')]), int(r[hdr.index('# Samples')])
        except ValueError:
            continue
        k = (cur.split('/')[-1], int(r[0]))
        agg[k][0] += inst
        agg[k][1] += samp
        agg[k][2] = r[1]
tot = sum(v[0] for v in agg.values())
ts = max(1, sum(v[1] for v in agg.values()))
print('total warp instructions %d%s' % (tot, ('  = %.1f thread-instr / element' % (tot * 32 / nelem)) if nelem else ''))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    per = (' %5.1f/elem' % (v[0] * 32 / nelem)) if nelem else ''
    print('%-20s %4d %5.2f%%%s  samples %5.2f%% | %s' % (k[0], k[1], 100 * v[0] / tot, per, 100 * v[1] / ts, v[2].strip()[:100]))
