#!/bin/bash
# multi-GPU: bit-identity of sharded runs (columns and rows) and the bench line at N GPUs
set -x
N=${1:-2}
O=gpurun_out/r2e_n$N
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 tests/multigpu_check.py > $O/multigpu_check.txt 2>&1; echo "rc=$?" >> $O/multigpu_check.txt
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "rc=$?" >> $O/bench_n$N.err
timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 --impl reference > $O/bench_ref_n$N.json 2> $O/bench_ref_n$N.err
python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
tail -5 $O/multigpu_check.txt; cat $O/bench_n$N.json
