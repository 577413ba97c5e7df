#!/bin/bash
# where the sharded stage time goes: A/B with the development knobs (results void with a knob set)
mkdir -p gpurun_out
L=gpurun_out/r2m4.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/tools/sharded_peer.py"
: > $L
for tune in 0 1 3 7; do
  for cfg in "1000000 4" "100000 8"; do
    set -- $cfg
    echo "== tune $tune bodies $1 substeps $2" >> $L
    BEPUCUDA_TUNE=$tune,0,0,0 timeout 600 $T --bodies $1 --substeps $2 --iterations 2 --steps 10 >> $L 2>&1
  done
done
grep -E "^==|sharded over|rror" $L
