#!/bin/bash
set -x
O=gpurun_out/r2p
mkdir -p $O
DBG=$PWD/openopt_b200/_ab/liboode_dbg.so
for k in default rows1 rows2; do
  if [ $k = default ]; then unset OODE_TRIAL_KERNEL; else export OODE_TRIAL_KERNEL=$k; fi
  OODE_LIB_PATH=$DBG timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $O/bounds_$k.txt 2>&1; echo "rc=$?" >> $O/bounds_$k.txt
done
unset OODE_TRIAL_KERNEL
ls -la $O
