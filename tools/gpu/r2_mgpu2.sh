#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2m2.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/tools/sharded_peer.py"
echo "== bench 1M 4x2, 2 GPUs, phase timing" > $L
BEPUCUDA_TUNE=0,0,0,1 timeout 600 $T --bodies 1000000 --substeps 4 --iterations 2 --steps 10 >> $L 2>&1
echo "== stream mode" >> $L
BEPUCUDA_TUNE=0,0,0,1 timeout 600 $T --bodies 1000000 --substeps 4 --iterations 2 --steps 10 --stream >> $L 2>&1
grep -E "sharded over|bepucuda shard" $L
