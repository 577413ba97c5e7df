"""One constraint graph over N GPUs (bepucuda_set_boundary_bodies): every rank holds all bodies and its slab's share of every batch; after each
WarmStart / Solve stage the written body records are all-reduced with NCCL (int32 sum of bit patterns, one writer per body and stage).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/tools/sharded_nccl.py --bodies 20000 --check

--check compares every rank's final body state with the CPU oracle's single-process result, bit for bit (strict build). Prints ms per step."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np
import torch
import torch.distributed as dist

import bepuphysics2_b200 as bp
import util
from bepuphysics2_b200 import scenes
from bepuphysics2_b200.native import EXEC_STREAM

ap = argparse.ArgumentParser()
ap.add_argument("--bodies", type=int, default=100_000)
ap.add_argument("--frames", type=int, default=3)
ap.add_argument("--substeps", type=int, default=8)
ap.add_argument("--iterations", type=int, default=2)
ap.add_argument("--check", action="store_true")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"
dist.init_process_group("nccl", device_id=torch.device("cuda", local))

scene = scenes.shape_pile(args.bodies, seed=5)
kw = dict(substeps=args.substeps, velocity_iterations=args.iterations)
sim = util.make_sim(scene, **kw)
n = scene["bodies"].shape[0]
kept = 0
for tb in sim.type_batches():  # this rank's slab of every batch; other lanes become empty
    refs = tb.body_references
    first = refs[:, 0, :]
    owner = np.where(first >= 0, ((first & 0x3FFFFFFF).astype(np.int64) * world) // n, -1)
    mine = owner == rank
    refs[np.broadcast_to(~mine[:, None, :], refs.shape)] = -1
    kept += int(mine.sum())


class Raw:
    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<i4", "data": (ptr, False), "version": 3}


calls = [0]


def exchange(ptr, count, op, stream):
    t = torch.as_tensor(Raw(ptr, count), device="cuda")
    if os.environ.get("SHARDED_SYNC"):
        torch.cuda.synchronize()
        dist.all_reduce(t, op=dist.ReduceOp.SUM if op == 0 else dist.ReduceOp.MIN)
        torch.cuda.synchronize()
    else:
        with torch.cuda.stream(torch.cuda.ExternalStream(stream)):
            dist.all_reduce(t, op=dist.ReduceOp.SUM if op == 0 else dist.ReduceOp.MIN)
    calls[0] += 1
    return 0


ts = bp.CudaTimestepper(sim, device=local, strict_fp=args.check, execution_mode=EXEC_STREAM)
if world > 1:
    ts.set_exchange(exchange)
ts.describe()
ms = []
for f in range(args.frames):
    if f > 0:
        ts.refresh()
    ts.solve(1 / 60.0, download=True)
    ms.append(ts.timings().solve_ms)
    if args.check:
        ts.download_prestep()  # contact depths advance on the device (IncrementallyUpdateForSubstep); the oracle mutates them in place too
ts.close()
ok = True
if args.check:
    ref = util.make_sim(scene, **kw)
    want = util.run_oracle(ref, 1 / 60.0, frames=args.frames, threads=8, simd=True)["bodies"]
    cols = util.MOTION
    ok = bool(np.array_equal(want[:, cols].view(np.uint32), sim.bodies[:, cols].view(np.uint32)))
    if not ok:
        bad = np.flatnonzero((want[:, cols].view(np.uint32) != sim.bodies[:, cols].view(np.uint32)).any(axis=1))
        d = np.abs(want[:, cols].astype(np.float64) - sim.bodies[:, cols].astype(np.float64))
        print("rank %d: %d of %d bodies differ (first %s, last %s), max abs diff %.3e, slab boundary at %d" % (rank, bad.size, n, bad[:5], bad[-5:], d.max(), n // world), flush=True)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
t = torch.tensor([float(np.mean(ms[1:] if len(ms) > 1 else ms))], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("ranks %d  bodies %d  constraints/rank ~%d  exchanges/step %d  ms/step %.3f  %s" % (
        world, n, kept, calls[0] // max(1, args.frames), t.item(), ("bit-exact vs oracle on every rank" if flag.item() == 1 else "MISMATCH") if args.check else "(no check)"))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
