"""Summarise an ncu report: headline metrics + instruction mix (per element) + top stall lines.
usage: python scripts/ncu_summary.py report.ncu-rep [elements_per_launch]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
nelem = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__grid_size', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg.per_second',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_wait_per_warp_active.pct',
        'smsp__warp_issue_stalled_not_selected_per_warp_active.pct', 'smsp__warp_issue_stalled_barrier_per_warp_active.pct',
        'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct',
        'smsp__warp_issue_stalled_no_instruction_per_warp_active.pct', 'smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('=' * 100)
    for k in want:
        if k in d:
            print('%-75s %s %s' % (k, d[k], units[hdr.index(k)]))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = None
data = []
for r in rows:
    if r and r[0] == 'Address':
        if h is not None:
            break
        h = r
        continue
    if h and len(r) == len(h):
        data.append(dict(zip(h, r)))
tot = sum(int(d['Instructions Executed']) for d in data)
ts = sum(int(d['# Samples']) for d in data)
by, samp = collections.Counter(), collections.Counter()
for d in data:
    parts = d['Source'].split()
    op = parts[1] if parts[0].startswith('@') else parts[0]
    op = op.split('.')[0]
    by[op] += int(d['Instructions Executed'])
    samp[op] += int(d['# Samples'])
print('-' * 100)
print('total warp instructions %d%s' % (tot, ('  = %.1f thread-instr / element' % (tot * 32 / nelem)) if nelem else ''))
for op, c in by.most_common(24):
    print('%-10s %6.2f%% %s samples %5.1f%%' % (op, 100 * c / tot, ('per-elem %6.1f ' % (c * 32 / nelem)) if nelem else '', 100 * samp[op] / ts))
print('-' * 100)
print('top stall lines')
for d in sorted(data, key=lambda d: -int(d['# Samples']))[:14]:
    print('%5.1f%%  %s' % (100 * int(d['# Samples']) / ts, d['Source'][:90]))
