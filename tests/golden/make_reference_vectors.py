"""Generates tests/golden/reference_vectors.npz: known-answer vectors for every constraint function on the hot path, produced by the REFERENCE'S OWN
C# text (transpiled mechanically by oracle/ref_transpile/cs2cpp.py and run here, where /root/reference exists). The file is committed so that the
oracle can be held to the reference anywhere (the GPU box has no /root/reference); tests/test_oracle_pinned_to_reference.py consumes it.

Per constraint type id and stage (0 WarmStart, 1 Solve, 2 IncrementallyUpdateForSubstep): inputs (body states, velocities, prestep, accumulated
impulses) and the reference's outputs. Inputs: prestep rows from the seeded scene generators (valid data for each type), random body states.
Plus the PoseIntegration functions (orientation integration through the custom Sin/Cos, inertia rotation, both momentum-conserving updates) and
the constraint micro-benchmark inputs of DemoBenchmarks/{One,Two,Three,Four}BodyConstraintBenchmarks*.cs (unit inertia, identity orientation,
SpringSettings(20 pi, 2), 1000 x (WarmStart; Solve) at dt = 1/60), parsed from those files by this script.

    python tests/golden/make_reference_vectors.py
"""
import ctypes as C
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_transpile"))
import json  # noqa: E402

import build_ref  # noqa: E402

FP = C.POINTER(C.c_float)
DT = 1.0 / 240.0
SAMPLES = 6


def ptr(a):
    return a.ctypes.data_as(FP)


def load_ref():
    lib = C.CDLL(build_ref.build())
    lib.ref_eval_lane.argtypes = [C.c_int32, C.c_int32, FP, C.c_float, FP, FP, FP, C.c_int32]
    lib.ref_eval_integration.argtypes = [C.c_int32, FP, FP]
    return lib


# ---- constraint micro-benchmark inputs, parsed from the reference's benchmark sources ---------------------------------------------------------
def _split_top(s):
    parts, depth, cur = [], 0, []
    for ch in s:
        if ch in "({<":
            depth += 1
        elif ch in ")}>":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    if "".join(cur).strip():
        parts.append("".join(cur))
    return [p.strip() for p in parts]


def _scalar(expr):
    expr = expr.strip()
    expr = re.sub(r"MathF\.PI", repr(float(np.float32(np.pi))), expr)
    expr = re.sub(r"(?<![\w.])(\d+(?:\.\d+)?)f\b", r"\1", expr)
    expr = expr.replace("float.MaxValue", repr(float(np.finfo(np.float32).max)))
    # evaluate in float32 like the C# constant folding of float expressions
    val = eval(re.sub(r"(?<![\w.])(\d+(?:\.\d+)?(?:e[-+]?\d+)?)", r"np.float32(\1)", expr), {"np": np})  # noqa: S307 - arithmetic on literals parsed from the benchmark source
    return float(np.float32(val))


def _assign(prefix, expr, out):
    """Flattens one C# initializer expression into {field path: float}."""
    expr = expr.strip()
    m = re.match(r"^new(?:\s+[\w<>]+)?\s*(?:\(\s*\))?\s*\{(.*)\}$", expr, flags=re.S)
    if m:
        for part in _split_top(m.group(1)):
            k, v = part.split("=", 1)
            _assign(prefix + [k.strip()], v, out)
        return
    m = re.match(r"^Vector3Wide\.Broadcast\(\s*(?:new\s+Vector3\((.*)\)|Vector3\.(\w+))\s*\)$", expr, flags=re.S)
    if m:
        if m.group(2):
            vals = {"Zero": (0, 0, 0), "One": (1, 1, 1), "UnitX": (1, 0, 0), "UnitY": (0, 1, 0), "UnitZ": (0, 0, 1)}[m.group(2)]
        else:
            vals = [_scalar(x) for x in _split_top(m.group(1))]
            vals = vals * 3 if len(vals) == 1 else vals
        for axis, v in zip("XYZ", vals):
            out[".".join(prefix + [axis])] = float(v)
        return
    m = re.match(r"^Vector3Wide\.Normalize\(Vector3Wide\.Broadcast\(\s*new\s+Vector3\((.*)\)\s*\)\)$", expr, flags=re.S)
    if m:
        v = np.array([_scalar(x) for x in _split_top(m.group(1))], dtype=np.float32)
        v = v / np.sqrt((v * v).sum(dtype=np.float32))
        for axis, x in zip("XYZ", v):
            out[".".join(prefix + [axis])] = float(x)
        return
    m = re.match(r"^QuaternionWide\.Broadcast\(\s*Quaternion\.Identity\s*\)$", expr)
    if m:
        for axis, v in zip("XYZW", (0, 0, 0, 1)):
            out[".".join(prefix + [axis])] = float(v)
        return
    m = re.match(r"^new\s+Vector<float>\((.*)\)$", expr, flags=re.S)
    if m:
        out[".".join(prefix)] = _scalar(m.group(1))
        return
    m = re.match(r"^Vector<float>\.(Zero|One)$", expr)
    if m:
        out[".".join(prefix)] = 0.0 if m.group(1) == "Zero" else 1.0
        return
    raise ValueError("unparsed initializer: " + expr[:80])


def benchmark_inputs(reference, layouts):
    """(benchmark name, type id, prestep rows, positions of the bodies) for every benchmark method whose prestep initializer this parser understands."""
    by_struct = {t["prestep_struct"]: (int(tid), t) for tid, t in layouts["types"].items()}
    found = []
    for fname in sorted(os.listdir(os.path.join(reference, "DemoBenchmarks"))):
        if not re.match(r"^(One|Two|Three|Four)BodyConstraintBenchmarks(Deep)?\.cs$", fname):
            continue
        src = open(os.path.join(reference, "DemoBenchmarks", fname), encoding="utf-8-sig").read()
        for m in re.finditer(r"\[Benchmark\]\s*public\s+[^\n]*?\b(\w+)\(\)\s*\{", src):
            start = m.end()
            pm = re.compile(r"var\s+prestep\s*=\s*new\s+(\w+)\s*\{").search(src, start)
            if not pm or pm.start() - start > 200:
                continue
            depth, i = 0, pm.end() - 1
            while True:
                depth += src[i] == "{"
                depth -= src[i] == "}"
                if depth == 0:
                    break
                i += 1
            struct = pm.group(1)
            if struct not in by_struct:
                continue
            tid, t = by_struct[struct]
            fields = {}
            try:
                _assign([], "new " + struct + " " + src[pm.end() - 1:i + 1], fields)
            except (ValueError, KeyError, SyntaxError):
                continue
            unknown = set(fields) - set(t["prestep_rows"])
            if unknown:
                continue
            rows = np.array([fields.get(r, 0.0) for r in t["prestep_rows"]], dtype=np.float32)
            call = src[i:src.index("\n    }", i)]
            positions = [(0.0, 0.0, 0.0)] * t["bodies"]
            pos = re.findall(r"(new Vector3Wide\(\)|Vector3Wide\.Broadcast\(new Vector3\(([^)]*)\)\))\s*,\s*orientation", call)
            if len(pos) == t["bodies"]:
                positions = [tuple(_scalar(x) for x in p[1].split(",")) if p[1] else (0.0, 0.0, 0.0) for p in pos]
            found.append((fname[:-3] + "." + m.group(1), tid, rows, np.array(positions, dtype=np.float32)))
    return found


def main():
    from bepuphysics2_b200 import scenes
    from oracle import binding as ob

    lib = load_ref()
    layouts = json.load(open(os.path.join(ROOT, "tests", "golden", "type_layouts.json")))
    samples = {}
    for scene in (scenes.shape_pile(1500, seed=21, nonconvex_fraction=0.5), scenes.joint_zoo(600, per_type=24, seed=21), scenes.ragdolls(6, seed=21), scenes.ragdolls(6, seed=22, motor="servo")):
        for type_id, _, prestep in scene["constraints"]:
            samples.setdefault(int(type_id), []).append(np.asarray(prestep, dtype=np.float32)[:SAMPLES])
    rng = np.random.default_rng(2024)
    out = {}
    for type_id in sorted(samples):
        bodies, prestep_rows, impulse_rows = ob.type_info(type_id)
        pre = np.concatenate(samples[type_id])[:SAMPLES]
        n = pre.shape[0]
        states = np.zeros((n, bodies, 14), dtype=np.float32)
        states[:, :, 0:3] = rng.normal(0, 2, (n, bodies, 3))
        q = rng.normal(0, 1, (n, bodies, 4))
        states[:, :, 3:7] = q / np.linalg.norm(q, axis=2, keepdims=True)
        for i in range(n):
            for b in range(bodies):
                r = np.linalg.qr(rng.normal(0, 1, (3, 3)))[0]
                mm = r @ np.diag(rng.uniform(0.5, 4.0, 3)) @ r.T
                states[i, b, 7:13] = (mm[0, 0], mm[1, 0], mm[1, 1], mm[2, 0], mm[2, 1], mm[2, 2])
        states[:, :, 13] = rng.uniform(0.3, 2.0, (n, bodies))
        vel = rng.normal(0, 1.5, (n, bodies, 6)).astype(np.float32)
        imp = np.abs(rng.normal(0, 0.2, (n, impulse_rows))).astype(np.float32)
        res = {s: (np.zeros_like(vel), np.zeros_like(imp), np.zeros_like(pre)) for s in (0, 1, 2)}
        for i in range(n):
            for stage in (0, 1, 2):
                v, a, p = vel[i].copy(), imp[i].copy(), pre[i].copy()
                st = np.ascontiguousarray(states[i])
                assert lib.ref_eval_lane(type_id, stage, ptr(st), DT, ptr(p), ptr(a), ptr(v), 1) == 0
                res[stage][0][i], res[stage][1][i], res[stage][2][i] = v, a, p
        key = "lane_%02d_" % type_id
        out[key + "states"], out[key + "velocities"], out[key + "impulses"], out[key + "prestep"] = states, vel, imp, pre
        for stage in (0, 1, 2):
            out[key + "out%d_velocities" % stage], out[key + "out%d_impulses" % stage], out[key + "out%d_prestep" % stage] = res[stage]
    # PoseIntegration
    ops = {0: (8, 4), 1: (10, 6), 2: (19, 3), 3: (14, 3)}
    for op, (n_in, n_out) in ops.items():
        ins = rng.normal(0, 1, (16, n_in)).astype(np.float32)
        for row in ins:
            qpos = 6 if op == 1 else 0
            row[qpos:qpos + 4] /= np.linalg.norm(row[qpos:qpos + 4])
            if op == 0:
                row[7] = abs(row[7]) * 0.01
            if op in (2, 3):  # symmetric positive definite local inverse inertia (and world inertia for op 2)
                for at in ((4, 10) if op == 2 else (4,)):
                    r = np.linalg.qr(rng.normal(0, 1, (3, 3)))[0]
                    mm = r @ np.diag(rng.uniform(0.5, 4.0, 3)) @ r.T
                    row[at:at + 6] = (mm[0, 0], mm[1, 0], mm[1, 1], mm[2, 0], mm[2, 1], mm[2, 2])
                if op == 3:
                    row[13] = 1.0 / 240.0
        ins[0, 4:7] = 0 if op == 0 else ins[0, 4:7]  # zero angular velocity: the identity fallback of Integrate
        outs = np.zeros((16, n_out), dtype=np.float32)
        for i in range(16):
            assert lib.ref_eval_integration(op, ptr(ins[i]), ptr(outs[i])) == 0
        out["integration_%d_in" % op], out["integration_%d_out" % op] = ins, outs
    # constraint micro-benchmarks: 1000 x (WarmStart; Solve), dt = 1/60, from zero velocities and impulses
    names = []
    for name, tid, rows, positions in benchmark_inputs(build_ref.REFERENCE, layouts):
        bodies, prestep_rows, impulse_rows = ob.type_info(tid)
        st = np.zeros((bodies, 14), dtype=np.float32)
        st[:, 0:3] = positions
        st[:, 6] = 1
        st[:, 7] = st[:, 9] = st[:, 12] = st[:, 13] = 1
        v, a, p = np.zeros((bodies, 6), dtype=np.float32), np.zeros(impulse_rows, dtype=np.float32), rows.copy()
        for _ in range(1000):
            lib.ref_eval_lane(tid, 0, ptr(st), 1.0 / 60.0, ptr(p), ptr(a), ptr(v), 1)
            lib.ref_eval_lane(tid, 1, ptr(st), 1.0 / 60.0, ptr(p), ptr(a), ptr(v), 1)
        k = "bench_%02d_" % len(names)
        out[k + "type"], out[k + "prestep"], out[k + "states"], out[k + "out_velocities"], out[k + "out_impulses"] = np.int32(tid), rows, st, v, a
        names.append(name)
    out["bench_names"] = np.array(names)
    path = os.path.join(ROOT, "tests", "golden", "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote %s: %d constraint types x 3 stages x %d samples, 4 integration functions, %d benchmark chains (%s)" % (path, len(samples), SAMPLES, len(names), ", ".join(names)))


if __name__ == "__main__":
    main()
