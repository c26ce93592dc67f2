"""Record the DRAM traffic of a workload's trial kernel from an `ncu --set full` report into profiles/traffic.json,
stamped with the git sha and the hash of the kernel sources it was captured from (bench.py refuses a stale entry).

    python scripts/update_traffic.py WORKLOAD report.ncu-rep [kernel-name-substring]
"""
import csv
import datetime
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

name, rep = sys.argv[1], sys.argv[2]
want = sys.argv[3] if len(sys.argv) > 3 else 'de_trial'
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if want not in d['Kernel Name']:
        continue
    tot = 0.0
    for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        tot += float(d[k].replace(',', '')) * scale[units[hdr.index(k)]]
    path = os.path.join(ROOT, 'profiles', 'traffic.json')
    try:
        cur = json.load(open(path))
    except Exception:
        cur = {}
    sha = subprocess.run(['git', '-C', ROOT, 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True).stdout.strip()
    cur[name] = {'dram_bytes_per_launch': tot, 'kernel': d['Kernel Name'], 'report': os.path.relpath(rep, ROOT),
                 'git_sha': sha + ('+dirty' if subprocess.run(['git', '-C', ROOT, 'status', '--porcelain', 'openopt_b200/csrc'],
                                                               capture_output=True, text=True).stdout.strip() else ''),
                 'csrc_sha256': bench.csrc_sha256(), 'captured': datetime.date.today().isoformat(),
                 'gpu_time_ms': d.get('gpu__time_duration.sum')}
    json.dump(cur, open(path, 'w'), indent=1)
    print(name, cur[name])
    break
else:
    sys.exit('no kernel matching %r in %s' % (want, rep))
