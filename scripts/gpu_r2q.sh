#!/bin/bash
# final 8-GPU record: bit-identity, C3 bench at N = 8 / 4 / 2 / 1 on ONE box, C5 row-shard points
set -x
O=gpurun_out/r2q
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29511 tests/multigpu_check.py --p2p-only > $O/multigpu_check_n8.txt 2>&1; echo "rc=$?" >> $O/multigpu_check_n8.txt
timeout 300 python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 300 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
timeout 300 $TR --nproc-per-node 4 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 3 > $O/bench_n4.json 2> $O/bench_n4.err
timeout 300 $TR --nproc-per-node 8 --master-port 29514 bench.py --gpus 8 --steps 20 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err
timeout 300 $TR --nproc-per-node 8 --master-port 29515 bench.py --gpus 8 --steps 20 --warmup 3 > $O/bench_n8_b.json 2> $O/bench_n8_b.err
timeout 300 $TR --nproc-per-node 8 --master-port 29516 scripts/sweep_c5.py --no-cpu --min-log2 20 > $O/sweep_c5_n8.jsonl 2> $O/sweep_c5_n8.err
timeout 300 $TR --nproc-per-node 4 --master-port 29517 scripts/sweep_c5.py --no-cpu --min-log2 20 > $O/sweep_c5_n4.jsonl 2> $O/sweep_c5_n4.err
timeout 300 python scripts/sweep_c5.py --no-cpu --min-log2 20 > $O/sweep_c5_n1.jsonl 2> $O/sweep_c5_n1.err
for f in n1 n2 n4 n8 n8_b; do python - <<PY
import json
try:
    d=json.load(open('$O/bench_$f.json')); r=d['roofline']
    print("$f", "ms/step %.4f"%d["ms_per_step"], "value %.3e"%d["value"], "trial %.4f sel %.4f"%(r["kernel_ms"], r["select_kernel_ms"]), "parity", (d.get("parity_vs_1gpu") or {}).get("records"), "e2e %.3e"%d["e2e"]["value"]); [print("   rank", k, {a: round(b,4) for a,b in pr.items()}) for k,pr in enumerate(d.get("per_rank") or [])]
except Exception as e: print('$f ERR', e)
PY
done
grep -h "^{" $O/sweep_c5_n*.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('sweep NP=%9d N=%d %8.4f ms %.3e evals/s' % (d['population'], d['n_gpus'], d['ms_per_generation'], d['evals_per_s']))"
tail -3 $O/multigpu_check_n8.txt; grep -c True $O/multigpu_check_n8.txt
