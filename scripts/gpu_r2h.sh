#!/bin/bash
set -x
O=gpurun_out/r2h
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.txt 2>&1; echo "rc=$?" >> $O/pytest.txt
for k in rows1 rows2 quad; do
  OODE_TRIAL_KERNEL=$k timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_expr_objectives.py -m gpu -q > $O/pytest_$k.txt 2>&1; echo "rc=$?" >> $O/pytest_$k.txt
done
python bench.py --workload c4 --steps 20 --warmup 3 > $O/bench_c4.json 2> $O/bench_c4.err
python bench.py --workload c2 --steps 200 --warmup 10 > $O/bench_c2.json 2> $O/bench_c2.err
python bench.py --workload c5_1m --steps 20 --warmup 3 > $O/bench_c5_1m.json 2> $O/bench_c5_1m.err
python bench.py --workload c5_16m --steps 10 --warmup 3 > $O/bench_c5_16m.json 2> $O/bench_c5_16m.err
python scripts/ab.py run tree --repeat 1 --shapes "griewank:200:262144 griewank:1000:262144 rastrigin:100:4194304 ackley:120:1000000" >> $O/ab.txt 2>&1
python examples/glp_1.py > $O/example_glp_1.txt 2>&1
python examples/glp_Ab_c.py > $O/example_glp_Ab_c.txt 2>&1
ls -la $O
