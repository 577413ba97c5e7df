#!/bin/bash
# Round 2, 2-GPU call: peer-sharded solve, parity check then timing.
mkdir -p gpurun_out
L=gpurun_out/r2m1.log
nvidia-smi topo -m > gpurun_out/r2m1_topo.txt 2>&1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/tools/sharded_peer.py"
echo "== check 20k graph" > $L
timeout 300 $T --check --bodies 20000 >> $L 2>&1
echo "== check 100k graph 8x2" >> $L
timeout 300 $T --check --bodies 100000 --substeps 8 --frames 2 >> $L 2>&1
echo "== bench 1M 4x2, 2 GPUs" >> $L
timeout 600 $T --bodies 1000000 --substeps 4 --iterations 2 --steps 20 >> $L 2>&1
echo "== bench 1M 4x2, 1 GPU (same tool)" >> $L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29512 tests/tools/sharded_peer.py --bodies 1000000 --substeps 4 --iterations 2 --steps 20 >> $L 2>&1

grep -v "^W0\|^\*\*\*\|OMP_NUM\|^$" $L | tail -40
