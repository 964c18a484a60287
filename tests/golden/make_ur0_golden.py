"""Displacement-triggered list rebuilds (DDC updateRate = 0): traces of the UNMODIFIED reference CPU path
(oracle/_ref/ref_dump) on the small golden decks with `updateRate=0;`.

Run in the build container after make_golden.py:  python tests/golden/make_ur0_golden.py
Writes tests/golden/ur0.npz: per deck the [NSTEPS, 16] trace (loop, eion, rk, ..., npairs, lastUpdate) and the
final positions.
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from refdump import read_records  # noqa: E402

REF_DUMP = os.path.join(ROOT, "oracle", "_ref", "ref_dump")
NSTEPS = 120
DECKS = ("waterbox", "popc_small", "ras_small")


def ur0_copy(deck, tmp):
    """copy of a golden deck with DDC updateRate = 0"""
    dst = os.path.join(tmp, deck)
    shutil.copytree(os.path.join(HERE, deck), dst, symlinks=True)
    p = os.path.join(dst, "object.data")
    s = open(p).read()
    assert "updateRate=20;" in s
    open(p, "w").write(s.replace("updateRate=20;", "updateRate=0;"))
    return dst


def run(deck):
    tmp = tempfile.mkdtemp(prefix="ur0_")
    dst = ur0_copy(deck, tmp)
    out = os.path.join(dst, "_ur0.bin")
    subprocess.check_call([REF_DUMP, out, str(NSTEPS), "0"], cwd=dst, stdout=open(os.path.join(dst, "_ur0.log"), "w"), stderr=subprocess.STDOUT)
    r = read_records(out)
    shutil.rmtree(tmp)
    return r


if __name__ == "__main__":
    out = {}
    for d in DECKS:
        r = run(d)
        tr = r["trace"].reshape(-1, 16)
        out[d + "_trace"] = tr
        for k in ("rx", "ry", "rz"):
            out[d + "_" + k] = r["sN_" + k]
        print(d, "rebuild loops:", sorted(set(int(x) for x in tr[:, 15])))
    np.savez_compressed(os.path.join(HERE, "ur0.npz"), **out)
