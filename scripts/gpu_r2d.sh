#!/bin/bash
set -x
O=gpurun_out/r2d
mkdir -p $O
for k in rows1 rows2 rows4; do
  OODE_TRIAL_KERNEL=$k timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $O/pytest_$k.txt 2>&1; echo "rc=$?" >> $O/pytest_$k.txt
done
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_default.txt 2>&1; echo "rc=$?" >> $O/pytest_default.txt
NARROW="ackley:120:1000000 ackley:128:1000000 rastrigin:100:4194304 sphere:100:4194304 ackley:256:1000000 rastrigin:100:10000"
for k in rows1 rows2 rows4; do
  python scripts/ab.py run tree --repeat 1 --shapes "$NARROW" --env OODE_TRIAL_KERNEL=$k >> $O/ab_narrow.txt 2>&1
done
W="ackley:1000:1000000 sphere:1000:1000000"
python scripts/ab.py run tree --repeat 2 --shapes "$W" --env OODE_TRIAL_KERNEL=rows1 >> $O/ab_wide.txt 2>&1
python scripts/ab.py run tree --repeat 2 --shapes "$W" --env OODE_TRIAL_KERNEL=rows2 >> $O/ab_wide.txt 2>&1
python bench.py --workload c4 --steps 20 --warmup 3 > $O/bench_c4.json 2> $O/bench_c4.err
ls -la $O
