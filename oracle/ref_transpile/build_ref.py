"""ORACLE PINNING — test infrastructure only. Recipe for oracle/_ref/libbepu_ref.so: the reference's own constraint / wide-math / pose-integration
C# sources, read where they lie under /root/reference, transpiled mechanically to C++ (cs2cpp.py) and compiled with g++ -ffp-contract=off.
Outputs go to oracle/_ref/ only (git-ignored; the built .so travels to the GPU box with the snapshot). Returns None when the reference tree is
absent and no prebuilt library exists."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "_ref")
LIB = os.path.join(OUT, "libbepu_ref.so")
REFERENCE = os.environ.get("BEPU_REFERENCE_ROOT", "/root/reference")


def build(force=False):
    have_reference = os.path.isdir(os.path.join(REFERENCE, "BepuPhysics", "Constraints"))
    sources = [os.path.join(HERE, f) for f in ("cs2cpp.py", "ref_runtime.h", "build_ref.py")]
    if os.path.exists(LIB) and not force and (not have_reference or all(os.path.getmtime(s) <= os.path.getmtime(LIB) for s in sources)):
        return LIB
    if not have_reference:
        return None
    os.makedirs(OUT, exist_ok=True)
    shutil.copy(os.path.join(HERE, "ref_runtime.h"), os.path.join(OUT, "ref_runtime.h"))
    header, harness = os.path.join(OUT, "bepu_ref_generated.h"), os.path.join(OUT, "bepu_ref_harness.cpp")
    subprocess.check_call([sys.executable, os.path.join(HERE, "cs2cpp.py"), REFERENCE, header, "--harness", harness])
    # -O1: the file is 10 k lines of inlined templates; -ffp-contract=off -fno-fast-math: RyuJIT never contracts or reassociates Vector<float> code
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-march=x86-64-v3", "-shared", "-o", LIB, harness])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
