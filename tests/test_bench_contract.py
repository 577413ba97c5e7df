"""The parts of bench.py's contract that can be held on a machine without a GPU: the reference arm prints exactly one JSON line with the agreed keys
(under torchrun only rank 0 prints), and our arm refuses to run without a CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)


def test_reference_arm_prints_one_json_line_with_the_contract_keys(libs):
    r = _run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--bodies", "3000", "--substeps", "2", "--iterations", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["steps"] == 2 and line["warmup"] == 1 and line["higher_is_better"] is True
    assert line["metric"].startswith("constraint-iterations/sec") and line["unit"] == "constraint-iterations/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "3000 bodies" in line["config"]["workload"] and line["config"]["substeps"] == 2


def test_reference_arm_is_silent_on_ranks_other_than_zero(libs):
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--bodies", "500", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_has_no_cpu_fallback(libs):
    import torch

    if torch.cuda.is_available():
        return  # on a GPU box the -m gpu suite and the bench itself cover this arm
    r = _run(["--steps", "1", "--warmup", "1", "--bodies", "500"])
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
