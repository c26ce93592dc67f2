#!/bin/bash
# galileo on one B200: parity tests, the bench line, a launch list and one `ncu --set full` capture of the children kernel
# (profiles/r02x_*, r02w_launches_*).  Run as: gpurun --timeout 400 -- 'bash scripts/gpu_r2x.sh'
set -x
O=gpurun_out/r2x
mkdir -p $O
timeout 240 python -m pytest tests/test_gpu_galileo.py -x -q -m gpu > $O/ga_tests.log 2>&1; echo "rc=$?" >> $O/ga_tests.log
timeout 120 python bench.py --workload ga_c5_1m --steps 10 --warmup 3 > $O/bench_ga_c5_1m.json 2> $O/bench_ga.err
# profiler passes only after the plain run has exited 0; numbers printed under ncu are never bench values
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/launches_ga.csv \
    python bench.py --workload ga_c5_1m --steps 2 --warmup 1 > $O/ncu_ga.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:ga_children_kernel -s 3 -c 1 -o $O/ga_children \
    python bench.py --workload ga_c5_1m --steps 2 --warmup 1 > $O/ncu_full.log 2>&1
ls -la $O
