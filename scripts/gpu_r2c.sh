#!/bin/bash
set -x
O=gpurun_out/r2c
mkdir -p $O
for k in rows1 rows2; do
  OODE_TRIAL_KERNEL=$k timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $O/pytest_$k.txt 2>&1; echo "rc=$?" >> $O/pytest_$k.txt
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $O/pytest_default.txt 2>&1; echo "rc=$?" >> $O/pytest_default.txt
NARROW="ackley:120:1000000 ackley:128:1000000 rastrigin:100:4194304 sphere:100:4194304 sphere:120:1000000"
for a in 2 16 32 64 128; do
  python scripts/ab.py run tree --repeat 1 --shapes "$NARROW" --env OODE_TRIAL_KERNEL=rows1 --env OODE_LD_ALIGN=$a >> $O/ab_ld.txt 2>&1
done
python scripts/ab.py run tree --repeat 1 --shapes "$NARROW" --env OODE_TRIAL_KERNEL=rows2 --env OODE_LD_ALIGN=128 >> $O/ab_ld.txt 2>&1
python scripts/ab.py run tree --repeat 1 --shapes "$NARROW" --env OODE_TRIAL_KERNEL=rows4 --env OODE_LD_ALIGN=128 >> $O/ab_ld.txt 2>&1
W="ackley:1000:1000000 sphere:1000:1000000"
python scripts/ab.py run tree --repeat 2 --shapes "$W" --env OODE_TRIAL_KERNEL=rows1 >> $O/ab_wide.txt 2>&1
python scripts/ab.py run tree --repeat 2 --shapes "$W" --env OODE_TRIAL_KERNEL=rows2 >> $O/ab_wide.txt 2>&1
python scripts/ab.py run tree --repeat 1 --shapes "$W" --env OODE_TRIAL_KERNEL=rows2 --env OODE_ROWS_WARPS=10 >> $O/ab_wide.txt 2>&1
python scripts/ab.py run tree --repeat 1 --shapes "$W" --env OODE_TRIAL_KERNEL=rows2 --env OODE_ROWS_WARPS=8 >> $O/ab_wide.txt 2>&1
python scripts/ab.py run tree --repeat 1 --shapes "$W" --env OODE_TRIAL_KERNEL=rows4 >> $O/ab_wide.txt 2>&1
python scripts/ab.py run tree --repeat 1 --shapes "$W" --env OODE_TRIAL_KERNEL=rows1 --env OODE_LD_ALIGN=128 >> $O/ab_wide.txt 2>&1
ls -la $O
