#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2m5.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/tools/sharded_peer.py"
echo "== check 20k fused" > $L
timeout 600 $T --bodies 20000 --check >> $L 2>&1
echo "== check 100k 8x2 fused" >> $L
timeout 900 $T --bodies 100000 --substeps 8 --iterations 2 --frames 3 --check >> $L 2>&1
for tune in 0 1 3; do
  for cfg in "1000000 4" "100000 8"; do
    set -- $cfg
    echo "== tune $tune bodies $1 substeps $2" >> $L
    BEPUCUDA_TUNE=$tune,0,0,0 timeout 600 $T --bodies $1 --substeps $2 --iterations 2 --steps 10 >> $L 2>&1
  done
done
echo "== 1M 8x2" >> $L
timeout 600 $T --bodies 1000000 --substeps 8 --iterations 2 --steps 10 >> $L 2>&1
grep -E "^==|sharded over|rror|differ" $L
