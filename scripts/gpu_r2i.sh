#!/bin/bash
set -x
O=gpurun_out/r2i
mkdir -p $O
SH="rastrigin:100:4194304 ackley:120:1000000 ackley:128:1000000 griewank:200:262144 ackley:256:1000000 ackley:1000:1000000"
for k in rows1 rows2; do for st in 2 3 4; do
  python scripts/ab.py run tree --repeat 1 --shapes "$SH" --env OODE_TRIAL_KERNEL=$k --env OODE_ROWS_STAGES=$st >> $O/ab_stages.txt 2>&1
done; done
python scripts/ab.py run tree --repeat 1 --shapes "$SH" >> $O/ab_stages.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_expr_objectives.py -m gpu -q > $O/pytest.txt 2>&1; echo "rc=$?" >> $O/pytest.txt
OODE_ROWS_STAGES=4 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q > $O/pytest_s4.txt 2>&1; echo "rc=$?" >> $O/pytest_s4.txt
cat $O/ab_stages.txt
