#!/bin/bash
# fused pushes: parity on 2 GPUs, then timing against the push-list variant and the 1-GPU plain solve
mkdir -p gpurun_out
L=gpurun_out/r2m3.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/tools/sharded_peer.py"
T1="python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29512 tests/tools/sharded_peer.py"
echo "== check 20k fused" > $L
timeout 600 $T --bodies 20000 --check >> $L 2>&1
echo "== check 100k 8x2 fused" >> $L
timeout 900 $T --bodies 100000 --substeps 8 --iterations 2 --frames 2 --check >> $L 2>&1
echo "== check 20k push lists" >> $L
timeout 600 $T --bodies 20000 --check --push-lists >> $L 2>&1
echo "== bench 1M 4x2, 2 GPUs fused" >> $L
timeout 600 $T --bodies 1000000 --substeps 4 --iterations 2 --steps 10 >> $L 2>&1
echo "== bench 1M 8x2, 2 GPUs fused" >> $L
timeout 600 $T --bodies 1000000 --substeps 8 --iterations 2 --steps 10 >> $L 2>&1
echo "== bench 1M 4x2, 1 GPU peer mode" >> $L
timeout 600 $T1 --bodies 1000000 --substeps 4 --iterations 2 --steps 10 >> $L 2>&1
echo "== bench 100k 8x2, 2 GPUs fused" >> $L
timeout 600 $T --bodies 100000 --substeps 8 --iterations 2 --steps 10 >> $L 2>&1
grep -E "^==|sharded over|bepucuda shard|differ|Error|error" $L
