#!/bin/bash
set -x
O=gpurun_out/r2g
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest.txt 2>&1; echo "rc=$?" >> $O/pytest.txt
./scripts/micro/gather_peak > $O/gather_peak.txt 2>&1
python bench.py > $O/bench_c3.json 2> $O/bench_c3.err
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:de_trial_rows -s 3 -c 1 -f -o $O/prof_c3_rows python bench.py --steps 2 --warmup 1 > $O/ncu_c3.log 2>&1
$NCU -k regex:de_trial_rows -s 5 -c 1 -f -o $O/prof_d120_rows python scripts/perf_probe.py ackley 120 1000000 4 32.768 > $O/ncu_d120.log 2>&1
$NCU -k regex:de_trial_rows -s 5 -c 1 -f -o $O/prof_d100_rows python scripts/perf_probe.py rastrigin 100 1048576 4 > $O/ncu_d100.log 2>&1
$NCU -k regex:de_trial_rows -s 3 -c 1 -f -o $O/prof_c4_rows python bench.py --workload c4 --steps 2 --warmup 1 > $O/ncu_c4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c3.csv python bench.py --steps 2 --warmup 1 > $O/ncu_launch.log 2>&1
for r in c3_rows:1000000000 d120_rows:120000000 d100_rows:104857600 c4_rows:52428800; do
  n=${r%%:*}; e=${r##*:}
  python scripts/ncu_summary.py $O/prof_$n.ncu-rep $e > $O/prof_${n}_summary.txt 2>&1
  python scripts/ncu_lines.py $O/prof_$n.ncu-rep $e 70 > $O/prof_${n}_lines.txt 2>&1
done
python scripts/update_traffic.py c3 $O/prof_c3_rows.ncu-rep de_trial_rows > $O/traffic_update.txt 2>&1
cp profiles/traffic.json $O/traffic.json
rm -f $O/prof_d120_rows.ncu-rep $O/prof_d100_rows.ncu-rep $O/prof_c4_rows.ncu-rep
ls -la $O
