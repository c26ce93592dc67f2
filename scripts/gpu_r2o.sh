#!/bin/bash
# bounds-assert build through the parity suite (every kernel form), then the product build: tests, bench, ncu capture
set -x
O=gpurun_out/r2o
mkdir -p $O
DBG=$PWD/openopt_b200/_ab/liboode_dbg.so
for k in default rows1 rows2 quad; do
  if [ $k = default ]; then unset OODE_TRIAL_KERNEL; else export OODE_TRIAL_KERNEL=$k; fi
  OODE_LIB_PATH=$DBG timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > $O/bounds_$k.txt 2>&1; echo "rc=$?" >> $O/bounds_$k.txt
done
unset OODE_TRIAL_KERNEL
OODE_LIB_PATH=$DBG timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -k "numpy" > $O/bounds_fullsize.txt 2>&1; echo "rc=$?" >> $O/bounds_fullsize.txt
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.txt 2>&1; echo "rc=$?" >> $O/pytest.txt
python bench.py > $O/bench_c3.json 2> $O/bench_c3.err
python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1
ls -la $O
