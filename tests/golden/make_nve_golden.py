"""10 000-step NVE energy traces of the UNMODIFIED reference CPU path (oracle/_ref/ref_dump) on the small golden
decks: the drift bar of BASELINE.json ("NVE drift over 10k steps no worse than the reference").

Run in the build container after make_golden.py:  python tests/golden/make_nve_golden.py
Writes tests/golden/nve10k.npz: for every deck an array [n_samples, 3] = (loop, eion, rk), sampled every 50 steps.
"""
import os
import subprocess
import sys
import tempfile
import shutil

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from refdump import read_records  # noqa: E402

REF_DUMP = os.path.join(ROOT, "oracle", "_ref", "ref_dump")
NSTEPS, EVERY = 10000, 50
DECKS = ("waterbox", "popc_small")


def run(deck):
    tmp = tempfile.mkdtemp(prefix="nve_" + deck)
    dst = os.path.join(tmp, deck)
    shutil.copytree(os.path.join(HERE, deck), dst, symlinks=True)
    out = os.path.join(dst, "_nve.bin")
    subprocess.check_call([REF_DUMP, out, str(NSTEPS), "0", "light"], cwd=dst, stdout=open(os.path.join(dst, "_nve.log"), "w"),
                          stderr=subprocess.STDOUT)
    r = read_records(out)
    tr = r["trace"].reshape(-1, 16)
    e0 = r["s0_energy"]
    rows = [(0.0, e0[0], e0[1])] + [(tr[k, 0], tr[k, 1], tr[k, 2]) for k in range(EVERY - 1, NSTEPS, EVERY)]
    shutil.rmtree(tmp)
    return np.array(rows), int(r["nion"][0])


if __name__ == "__main__":
    out = {}
    for d in DECKS:
        out[d], n = run(d)
        e = out[d][:, 1] + out[d][:, 2]
        print(d, n, "beads: Etot(0) %.12g Etot(10k) %.12g drift/bead %.3e" % (e[0], e[-1], (e[-1] - e[0]) / n))
        out[d + "_n"] = np.array([n])
    np.savez_compressed(os.path.join(HERE, "nve10k.npz"), **out)
