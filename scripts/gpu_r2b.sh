#!/bin/bash
# round-2 second GPU call: parity of the new row kernel (all forms), A/B against the round-1 library
set -x
O=gpurun_out/r2b
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.txt 2>&1; echo "rc=$?" >> $O/pytest.txt
for k in rows1 rows2 rows4 quad; do
  OODE_TRIAL_KERNEL=$k timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q > $O/pytest_$k.txt 2>&1; echo "rc=$?" >> $O/pytest_$k.txt
done
SH="ackley:1000:1000000 ackley:120:1000000 ackley:128:1000000 rastrigin:100:4194304 rastrigin:100:10000 ackley:256:1000000 sphere:1000:1000000"
python scripts/ab.py run r1 tree --repeat 1 --shapes "$SH" > $O/ab.txt 2>&1
python scripts/ab.py run tree --repeat 1 --shapes "ackley:120:1000000 ackley:128:1000000 rastrigin:100:4194304 ackley:256:1000000" --env OODE_TRIAL_KERNEL=rows2 >> $O/ab.txt 2>&1
python scripts/ab.py run tree --repeat 1 --shapes "ackley:120:1000000 ackley:128:1000000 rastrigin:100:4194304 ackley:256:1000000 ackley:1000:1000000" --env OODE_TRIAL_KERNEL=rows1 >> $O/ab.txt 2>&1
python scripts/ab.py run tree --repeat 1 --shapes "ackley:1000:1000000" --env OODE_ROWS_WARPS=16 >> $O/ab.txt 2>&1
python scripts/ab.py run tree --repeat 1 --shapes "ackley:1000:1000000" --env OODE_TRIAL_KERNEL=rows2 >> $O/ab.txt 2>&1
python bench.py --workload c4 --steps 20 --warmup 3 > $O/bench_c4.json 2> $O/bench_c4.err
python bench.py --workload c2 --steps 200 --warmup 10 > $O/bench_c2.json 2> $O/bench_c2.err
python bench.py > $O/bench_c3.json 2> $O/bench_c3.err
ls -la $O
