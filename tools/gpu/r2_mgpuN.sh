#!/bin/bash
# usage: r2_mgpuN.sh N  -- parity + bench.py (islands + one graph sharded) on N GPUs
N=$1
mkdir -p gpurun_out
L=gpurun_out/r2mN_$N.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== check 100k 8x2, $N GPUs" > $L
timeout 900 $T tests/tools/sharded_peer.py --bodies 100000 --substeps 8 --iterations 2 --frames 2 --check >> $L 2>&1
echo "== bench.py --gpus $N" >> $L
timeout 900 $T bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_N$N.json 2>> $L
cat gpurun_out/bench_N$N.json >> $L
echo "== 1M 8x2, $N GPUs" >> $L
timeout 600 $T tests/tools/sharded_peer.py --bodies 1000000 --substeps 8 --iterations 2 --steps 10 >> $L 2>&1
grep -E "^==|sharded over|rror|differ" $L
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_N$N.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","n_gpus","ms_per_step","ms_per_step_per_rank")})
print(d.get("one_graph_sharded"))
PY
