#!/bin/bash
# usage: r2_mgpu_check_bench.sh N -- bit-exact check of the peer-sharded solve against the oracle, then bench.py (islands + one graph sharded) on N GPUs.
# (every step logs to its own file: NCCL re-opens /dev/stderr for writing, which truncates a shared append log)
N=$1
mkdir -p gpurun_out
P=gpurun_out/r2m${N}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $T tests/tools/sharded_peer.py --bodies 100000 --substeps 8 --iterations 2 --frames 2 --check > ${P}_check.log 2>&1
echo "check rc=$?" >> ${P}_check.log
timeout 900 $T bench.py --gpus $N --steps 10 --warmup 3 > ${P}_bench.json 2> ${P}_bench.err
echo "bench rc=$?" >> ${P}_bench.err
grep -E "sharded over|rror|differ|identical|check rc|MISMATCH|bit" ${P}_check.log | tail -8
tail -3 ${P}_bench.err
python - <<PY
import json
d=json.loads(open("${P}_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","n_gpus","ms_per_step","ms_per_step_per_rank")})
print(d.get("one_graph_sharded"))
PY
