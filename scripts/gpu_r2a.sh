#!/bin/bash
# round-2 first GPU call: state of the current kernels (profiles the verdict asked for), pin16 A/B, sanitizer pass
set -x
O=gpurun_out/r2a
mkdir -p $O
(nproc; free -g; nvidia-smi --query-gpu=name,clocks.max.sm,power.limit,memory.total --format=csv; nvidia-smi topo -m) > $O/box.txt 2>&1
python scripts/ab.py run tree pin16 --repeat 2 > $O/ab_pin16.txt 2>&1
python scripts/perf_probe.py > $O/probe.txt 2>&1
python bench.py --workload c4 --steps 20 --warmup 3 > $O/bench_c4.json 2> $O/bench_c4.err
python bench.py --workload c2 --steps 200 --warmup 10 > $O/bench_c2.json 2> $O/bench_c2.err
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:de_trial_tma2 -s 5 -c 1 -f -o $O/prof_d100_tma2 python scripts/perf_probe.py rastrigin 100 1048576 4 > $O/ncu_d100.log 2>&1
$NCU -k regex:de_trial_tma2 -s 5 -c 1 -f -o $O/prof_d120_tma2 python scripts/perf_probe.py ackley 120 1000000 4 32.768 > $O/ncu_d120.log 2>&1
$NCU -k regex:de_trial_kernel -s 3 -c 1 -f -o $O/prof_c4_quad python bench.py --workload c4 --steps 2 --warmup 1 > $O/ncu_c4.log 2>&1
$NCU -k regex:de_select -s 3 -c 1 -f -o $O/prof_select_c3 python scripts/perf_probe.py ackley 1000 1000000 3 32.768 > $O/ncu_sel.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_c2.csv python bench.py --workload c2 --steps 20 --warmup 3 > $O/ncu_c2.log 2>&1
for r in d100_tma2:104857600 d120_tma2:120000000 c4_quad:52428800; do
  n=${r%%:*}; e=${r##*:}
  python scripts/ncu_summary.py $O/prof_$n.ncu-rep $e > $O/prof_${n}_summary.txt 2>&1
  python scripts/ncu_lines.py $O/prof_$n.ncu-rep $e 60 > $O/prof_${n}_lines.txt 2>&1
done
python scripts/ncu_summary.py $O/prof_select_c3.ncu-rep 1000000 > $O/prof_select_c3_summary.txt 2>&1
rm -f $O/prof_select_c3.ncu-rep $O/prof_c4_quad.ncu-rep $O/prof_d120_tma2.ncu-rep
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --target-processes all python scripts/sanitize_probe.py > $O/sanitizer_$tool.log 2>&1
  echo "rc=$?" >> $O/sanitizer_$tool.log
done
ls -la $O
