#!/bin/bash
# 8-GPU box: bit-identity at N=8, bench at N=8 / 4 / 1 on the same box
set -x
O=gpurun_out/r2n
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"

timeout 300 $TR --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err
timeout 300 $TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 3 > $O/bench_n8_b.json 2> $O/bench_n8_b.err
timeout 300 $TR --nproc-per-node 4 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 3 > $O/bench_n4.json 2> $O/bench_n4.err
timeout 300 python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
for f in n8 n8_b n4 n1; do python - <<PY
import json
try:
    d=json.load(open('$O/bench_$f.json')); r=d['roofline']
    print("$f", "ms/step %.4f"%d["ms_per_step"], "value %.3e"%d["value"], "trial %.4f sel %.4f"%(r["kernel_ms"], r["select_kernel_ms"]), "parity", (d.get("parity_vs_1gpu") or {}).get("records"), "e2e %.3e"%d["e2e"]["value"]); [print("   rank", k, {a: round(b,4) for a,b in pr.items()}) for k,pr in enumerate(d.get("per_rank") or [])]
except Exception as e: print('$f ERR', e)
PY
done

