"""Regenerate the golden fixtures: decks (synthetic, fixed seeds, plus the reference's own
examples/waterbox with the NVE edits of SURVEY.md Appendix C) and the outputs of the
UNMODIFIED reference CPU path on them (oracle/_ref/ref_dump, built by oracle/build_ref.sh).

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
"""
import os
import re
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ddcmd_b200 import synth  # noqa: E402
from refdump import read_records  # noqa: E402

REF_DUMP = os.path.join(ROOT, "oracle", "_ref", "ref_dump")
NSTEPS, EVERY = 40, 20


def run_ref(deck_dir):
    out = os.path.join(deck_dir, "_ref.bin")
    subprocess.check_call([REF_DUMP, out, str(NSTEPS), str(EVERY)], cwd=deck_dir, stdout=open(os.path.join(deck_dir, "_ref.log"), "w"),
                          stderr=subprocess.STDOUT)
    rec = read_records(out)
    np.savez_compressed(os.path.join(deck_dir, "ref.npz"), **rec)
    for f in ("_ref.bin", "_ref.log", "ddd.data", "ddcMD.header", "hpm.data", "data"):
        p = os.path.join(deck_dir, f)
        if os.path.exists(p):
            os.remove(p)
    for d in os.listdir(deck_dir):
        if d.startswith("snapshot.") and d != "snapshot.mem":
            shutil.rmtree(os.path.join(deck_dir, d))


def waterbox():
    src = "/root/reference/examples/waterbox"
    dst = os.path.join(HERE, "waterbox")
    shutil.rmtree(dst, ignore_errors=True)
    os.makedirs(os.path.join(dst, "snapshot.mem"))
    for f in ("martini.data", "restraint.data", "snapshot.mem/atoms#000000", "snapshot.mem/restart"):
        shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
    obj = open(os.path.join(src, "object.data")).read()
    obj = obj.replace("deltaloop=10;", "deltaloop=500;").replace("printrate=1;", "printrate=50;")
    obj = re.sub(r"^nglf INTEGRATOR \{type = NGLFCONSTRAINT.*$", "nglf INTEGRATOR {type = NGLF; }", obj, flags=re.M)
    obj = re.sub(r"^group GROUP \{ type = LANGEVIN.*$", "group GROUP { type = FREE; }", obj, flags=re.M)
    obj = re.sub(r"^free GROUP \{ type = LANGEVIN.*$", "free GROUP { type = FREE; }", obj, flags=re.M)
    open(os.path.join(dst, "object.data"), "w").write(obj)
    os.symlink("snapshot.mem/restart", os.path.join(dst, "restart"))
    run_ref(dst)


def synthetic(name, builder, restraints=None):
    dst = os.path.join(HERE, name)
    shutil.rmtree(dst, ignore_errors=True)
    s = builder()
    rs = restraints(s) if restraints else None
    s.write_deck(dst, deltaloop=10, printrate=10, restraints=rs)
    run_ref(dst)


def ras_restraints(s):
    nat = np.array([r.natoms for r in s.residues])[s.mol_res][s.bead_mol]
    prot = np.nonzero(nat > 12)[0]
    po4 = np.nonzero((s.bead_atom == 1) & (nat == 12))[0]
    return [(int(i), 500.0, (1, 1, 1)) for i in prot[0:60:7]] + [(int(b), 200.0, (0, 0, 1)) for b in po4[:10]]


if __name__ == "__main__":
    subprocess.check_call([os.path.join(ROOT, "oracle", "build_ref.sh")])
    waterbox()
    synthetic("popc_small", lambda: synth.make_membrane(lx=57.0, ly=57.0, lz=112.0, seed=1))
    synthetic("ras_small", lambda: synth.make_membrane(lx=68.0, ly=57.0, lz=160.0, seed=2, protein_beads=120), ras_restraints)
    # two cells per axis in x and y (33.9 A box edge against the 15 A list radius): the deduplicated stencil of both list builds
    synthetic("tiny2", lambda: synth.make_membrane(lx=34.0, ly=34.0, lz=100.0, seed=5))
    print("golden fixtures written under", HERE)
