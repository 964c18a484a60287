"""Build tests/cpu_emu/libddcmd_b200_emu.so: the product's CUDA sources compiled with g++ against the CPU emulation
shim (tests/cpu_emu/shim).  TEST INFRASTRUCTURE ONLY - exercises kernel logic in the no-GPU container; the product
library never loads it and has no CPU path."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "ddcmd_b200", "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libddcmd_b200_emu.so")


def _deps():
    d = [os.path.join(HERE, "shim", f) for f in os.listdir(os.path.join(HERE, "shim"))]
    for root, _, files in os.walk(CSRC):
        d += [os.path.join(root, f) for f in files]
    d += [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    return d


def build(force=False, verbose=False):
    if not force and os.path.exists(LIB) and all(os.path.getmtime(f) <= os.path.getmtime(LIB) for f in _deps()):
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    objs = []
    for f in ("units.c", "objdb.c", "deck.c", "snapshot.c"):
        o = os.path.join(OBJ, f + ".o")
        subprocess.check_call(["gcc", "-std=gnu99", "-O2", "-ffp-contract=off", "-fPIC", "-w", "-c", os.path.join(CSRC, "host", f), "-o", o])
        objs.append(o)
    o = os.path.join(OBJ, "api.o")
    cmd = ["g++", "-x", "c++", "-std=c++17", "-O2", "-g", "-ffp-contract=off", "-fPIC", "-fno-omit-frame-pointer", "-DDDCB200_EMU_IMPL",
           "-I", os.path.join(HERE, "shim"), "-Wall", "-Wno-unused-function", "-Wno-unknown-pragmas", "-Wno-unused-variable",
           "-c", os.path.join(CSRC, "api.cu"), "-o", o]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    objs.append(o)
    subprocess.check_call(["g++", "-shared", "-o", LIB] + objs + ["-lm"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
