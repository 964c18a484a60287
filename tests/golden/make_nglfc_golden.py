"""NGLFCONSTRAINT goldens: traces of the UNMODIFIED reference CPU path
(oracle/_ref/ref_dump) on the golden decks with the INTEGRATOR switched to NGLFCONSTRAINT and/or the GROUPs to
LANGEVIN (tests/nglfc_decks.py makes the variants; `waterbox full` is the configuration the reference ships in
examples/waterbox/object.data with randomizeSeed=0).

Run in the build container after make_golden.py:  python tests/golden/make_nglfc_golden.py
Writes tests/golden/nglfc.npz: per case the [NSTEPS,16] trace, the box edges after every step, the final
rz / vz and the final per-bead LCG64 states; for one case also the default LCG64 tables at start-up.
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
import nglfc_decks  # noqa: E402

REF_DUMP = os.path.join(ROOT, "oracle", "_ref", "ref_dump")
NSTEPS = 40
CASES = [("waterbox", "full"), ("popc_small", "lang"), ("popc_small", "baro"), ("popc_small", "full"), ("ras_small", "lang"),
         ("ras_small", "full")]


def run(deck, variant):
    tmp = tempfile.mkdtemp(prefix="nglfc_")
    dst = nglfc_decks.make_variant(HERE, deck, variant, tmp)
    out = os.path.join(dst, "_o.bin")
    subprocess.check_call([REF_DUMP, out, str(NSTEPS), "0"], cwd=dst, stdout=open(os.path.join(dst, "_o.log"), "w"), stderr=subprocess.STDOUT)
    r = read_records(out)
    shutil.rmtree(tmp)
    return r


if __name__ == "__main__":
    out = {}
    for deck, variant in CASES:
        r = run(deck, variant)
        key = "%s_%s_" % (deck, variant)
        out[key + "trace"] = r["trace"].reshape(-1, 16)
        out[key + "box"] = r["boxtrace"].reshape(-1, 3)
        out[key + "rz"] = r["sN_rz"]
        out[key + "vz"] = r["sN_vz"]
        if nglfc_decks.VARIANTS[variant][0]:
            out[key + "rng"] = r["sN_rng_state"]
        if (deck, variant) == ("waterbox", "full"):
            out["waterbox_rng0_state"] = r["s0_rng_state"]
            out["waterbox_rng0_mp"] = r["s0_rng_mp"].reshape(-1, 2)
        print(deck, variant, "box", out[key + "box"][[0, -1]].tolist())
    np.savez_compressed(os.path.join(HERE, "nglfc.npz"), **out)
