#!/bin/bash
# 8-GPU experiment: exchange variants on one box
set -x
O=gpurun_out/r2r
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
timeout 300 python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 300 $TR --master-port 29514 bench.py --gpus 8 --steps 20 --warmup 3 > $O/bench_buffered.json 2> $O/bench_buffered.err
OODE_XCHG_UNBUFFERED=1 timeout 300 $TR --master-port 29515 bench.py --gpus 8 --steps 20 --warmup 3 > $O/bench_unbuffered.json 2> $O/bench_unbuffered.err
OODE_EXCHANGE=nccl timeout 300 $TR --master-port 29516 bench.py --gpus 8 --steps 20 --warmup 3 > $O/bench_nccl.json 2> $O/bench_nccl.err
OODE_DEBUG_LOCAL_STORES=1 timeout 300 $TR --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 3 > $O/bench_localstores.json 2> $O/bench_localstores.err
for f in n1 buffered unbuffered nccl localstores; do python - <<PY
import json
try:
    d=json.load(open('$O/bench_$f.json')); r=d['roofline']
    print("$f", "ms/step %.4f"%d["ms_per_step"], "value %.3e"%d["value"], "trial %.4f sel %.4f"%(r["kernel_ms"], r["select_kernel_ms"]), "parity", (d.get("parity_vs_1gpu") or {}).get("records")); [print("   rank", k, {a: round(b,4) for a,b in pr.items()}) for k,pr in enumerate(d.get("per_rank") or [])]
except Exception as e: print('$f ERR', e)
PY
done
