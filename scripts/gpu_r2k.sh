#!/bin/bash
set -x
N=${1:-2}
O=gpurun_out/r2k_n$N
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > $O/bench_p2p.json 2> $O/bench_p2p.err
OODE_EXCHANGE=nccl timeout 300 $TR --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 > $O/bench_nccl.json 2> $O/bench_nccl.err
python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
python scripts/ab.py run tree --repeat 2 --shapes "ackley:120:1000000 ackley:128:1000000 ackley:1000:1000000" > $O/ab.txt 2>&1
for f in p2p nccl n1; do python - <<PY
import json
try:
    d=json.load(open('$O/bench_$f.json')); r=d['roofline']
    print('$f', 'ms/step %.4f'%d['ms_per_step'], 'value %.3e'%d['value'], 'trial %.4f sel %.4f'%(r['kernel_ms'], r['select_kernel_ms']))
    for k,pr in enumerate(d.get('per_rank') or []): print('   rank', k, {a: round(b,4) for a,b in pr.items()})
except Exception as e: print('$f ERR', e)
PY
done
cat $O/ab.txt
