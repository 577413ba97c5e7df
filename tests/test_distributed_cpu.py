"""world_size-2 gloo test of the N>1 path's host logic: each rank owns an independent island (seed + rank), the aggregate metric is the SUM of
per-rank constraint-iterations over the MAX of per-rank times, exactly what bench.py reduces with NCCL on GPUs."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_island_sharding_reduction(libs, tmp_path):
    script = tmp_path / "rank.py"
    script.write_text(textwrap.dedent("""
        import os, sys, json
        sys.path.insert(0, %r)
        import torch, torch.distributed as dist
        import bepuphysics2_b200 as bp
        from bepuphysics2_b200 import scenes
        dist.init_process_group("gloo")
        rank = dist.get_rank()
        sim = bp.Simulation(substeps=2, velocity_iterations=2)
        scenes.build(scenes.shape_pile(500, seed=5 + rank), sim)
        ci = sim.constraint_count * 2 * 2
        t = torch.tensor([1.0 + rank], dtype=torch.float64)          # pretend per-rank time
        c = torch.tensor([float(ci)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        if rank == 0:
            print(json.dumps({"ci": c.item(), "t": t.item(), "mine": ci}))
        dist.destroy_process_group()
    """ % ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json

    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    r = json.loads(line)
    assert r["t"] == 2.0
    assert r["ci"] > r["mine"] > 0  # different seeds -> different islands, both counted
