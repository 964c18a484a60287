"""In-tree build of libddcmd_b200.so: host C (gcc, -ffp-contract=off) + CUDA (nvcc, sm_100a only).

The .so is written next to this file so it travels to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libddcmd_b200.so")
EXE = os.path.join(HERE, "ddcMD_b200")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall,-Wno-unused-function",
]
GCC_FLAGS = ["-std=gnu99", "-O2", "-ffp-contract=off", "-fPIC", "-Wall", "-Wno-unused-function"]


def _sources():
    cu = [os.path.join(CSRC, "api.cu")]
    c = [os.path.join(CSRC, "host", f) for f in ("units.c", "objdb.c", "deck.c", "snapshot.c")]
    deps = []
    for root, _, files in os.walk(CSRC):
        deps += [os.path.join(root, f) for f in files]
    deps += [os.path.join(HERE, "..", "include", f) for f in ("ddcmd_b200.h", "ddcmd_b200_host.h")]
    return cu, c, deps


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    _, _, deps = _sources()
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")
    os.makedirs(OBJ, exist_ok=True)
    cu, c, _ = _sources()
    objs = []
    for src in c:
        o = os.path.join(OBJ, os.path.basename(src) + ".o")
        cmd = ["gcc"] + GCC_FLAGS + ["-c", src, "-o", o]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        objs.append(o)
    for src in cu:
        o = os.path.join(OBJ, os.path.basename(src) + ".o")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", o]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        objs.append(o)
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-lcudart", "-lnccl", "-lm"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    # the stand-alone driver (main of the reference, src/ddcMD.c:66-88, for Martini decks): ddcMD_b200 [-o object.data] [-r restart] [-s simulate]
    cmd = ["gcc"] + GCC_FLAGS + [os.path.join(CSRC, "host", "main.c"), "-o", EXE, "-L", HERE, "-lddcmd_b200", "-Wl,-rpath,$ORIGIN", "-lm"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
