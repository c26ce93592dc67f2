#!/bin/bash
# final single-GPU validation of the frozen kernel sources: tests, smoke, bench, ncu capture for the traffic stamp
set -x
O=gpurun_out/r2t
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.txt 2>&1; echo "rc=$?" >> $O/pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1
python bench.py > $O/bench_c3.json 2> $O/bench_c3.err
ncu --set full --clock-control none --import-source on -k regex:de_trial_rows -s 3 -c 1 -f -o $O/prof_c3_rows python bench.py --steps 2 --warmup 1 > $O/ncu_c3.log 2>&1
python scripts/ncu_summary.py $O/prof_c3_rows.ncu-rep 1000000000 > $O/prof_c3_rows_summary.txt 2>&1
python scripts/ncu_lines.py $O/prof_c3_rows.ncu-rep 1000000000 60 > $O/prof_c3_rows_lines.txt 2>&1
python scripts/update_traffic.py c3 $O/prof_c3_rows.ncu-rep de_trial_rows > $O/traffic_update.txt 2>&1
cp profiles/traffic.json $O/traffic.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c3.csv python bench.py --steps 2 --warmup 1 > $O/ncu_launch.log 2>&1
python bench.py > $O/bench_c3_with_traffic.json 2> $O/bench_c3_b.err
rm -f $O/prof_c3_rows.ncu-rep
