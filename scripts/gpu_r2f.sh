#!/bin/bash
set -x
N=2
O=gpurun_out/r2f
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
OODE_EXCHANGE=nccl timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > $O/bench_n2_nccl.json 2> $O/bench_n2_nccl.err
OODE_DEBUG_LOCAL_STORES=1 timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 > $O/bench_n2_localstores.json 2> $O/bench_n2_localstores.err
timeout 600 $TR --master-port 29514 bench.py --gpus $N --steps 20 --warmup 3 > $O/bench_n2_p2p.json 2> $O/bench_n2_p2p.err
for f in nccl localstores p2p; do python - <<PY
import json
d=json.load(open('$O/bench_n2_$f.json')); r=d['roofline']
print('$f', 'ms/step %.4f'%d['ms_per_step'], 'trial %.4f sel %.4f'%(r['kernel_ms'], r['select_kernel_ms']), 'wait', d['wait_ms'], d['parity_vs_1gpu']['records'])
PY
done
