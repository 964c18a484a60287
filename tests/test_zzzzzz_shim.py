"""ddcMD itself with the library plugged in at its plug-in seam (integration/ddcmd_shim.c, built by oracle/build_ref.sh from the
reference's own objects): POTENTIAL MARTINI ->eval_potential = martiniB200 (mode 1: ddcMD's nglf integrates on the host) and
INTEGRATOR NGLF ->eval_integrator = nglfB200 (mode 2: whole steps on the device).  Start-up, ddcenergy's bookkeeping,
kinetic_terms, eval_energyInfo, the molecular pressure and printinfo remain the reference's; the `data` file of the run must be
the one the unmodified reference writes for the same deck (run side by side here)."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "ddcMD_ref")
SHIM = os.path.join(ROOT, "oracle", "_ref", "ddcMD_shim")
SHIM_EMU = os.path.join(ROOT, "oracle", "_ref", "ddcMD_shim_emu")


def run_deck(golden_dir, deck, tmp_path, exe, args, tag, variant=None):
    if variant:
        import nglfc_decks
        sub = os.path.join(str(tmp_path), tag)
        os.makedirs(sub)
        d = nglfc_decks.make_variant(golden_dir, deck, variant, sub)
    else:
        d = os.path.join(str(tmp_path), "%s_%s" % (deck, tag))
        shutil.copytree(os.path.join(golden_dir, deck), d, symlinks=True)
    p = os.path.join(d, "object.data")
    s = open(p).read()
    s = re.sub(r"deltaloop=\d+;", "deltaloop=25;", s)
    s = re.sub(r"printrate=\d+;", "printrate=5;", s)
    open(p, "w").write(s)
    r = subprocess.run([exe] + args, cwd=d, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return open(os.path.join(d, "data")).read().splitlines()


def check_shim(golden_dir, deck, tmp_path, exe, variant=None):
    ref = run_deck(golden_dir, deck, tmp_path, REF, [], "ref", variant)
    assert len(ref) == 7                                            # header + loops 0, 5, ..., 25 (across the rebuild at 20)
    # with the NGLFCONSTRAINT variants (Langevin groups, constraints, barostat) ddcMD's own integrator stays: mode 1 only
    for mode in (("1",) if variant else ("1", "2")):
        got = run_deck(golden_dir, deck, tmp_path, exe, [mode], "mode" + mode, variant)
        assert got[0] == ref[0] and len(got) == len(ref)
        for a, b in zip(got[1:], ref[1:]):
            fa, fb = a.split(), b.split()
            assert fa[0] == fb[0]
            va, vb = np.array(fa[1:], float), np.array(fb[1:], float)
            scale = max(abs(vb[1]), abs(vb[2]))
            assert np.all(np.abs(va[1:4] - vb[1:4]) <= 1e-9 * scale + 2e-12), (mode, a, b)     # Etotal, Ekin, Epot
            assert abs(va[4] - vb[4]) <= 1e-9 * abs(vb[4]) + 2e-8                              # temperature
            assert abs(va[5] - vb[5]) <= 1e-6 * max(abs(vb[5]), 100.0)                         # molecular pressure
            assert np.all(np.abs(va[6:] - vb[6:]) <= 1e-9 * np.abs(vb[6:]) + 2e-8)            # volume, box (barostat in the variants)


@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(SHIM_EMU)), reason="oracle/_ref binaries not built")
@pytest.mark.parametrize("deck", ["popc_small", "ras_small"])
def test_ddcmd_with_the_emulated_library_plugged_in(golden_dir, tmp_path, deck):
    """The seam in the build container: the shim binary linked against the CPU emulation of the kernels (test infrastructure)."""
    check_shim(golden_dir, deck, tmp_path, SHIM_EMU)


@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(SHIM_EMU)), reason="oracle/_ref binaries not built")
@pytest.mark.parametrize("deck", ["waterbox", "ras_small"])
def test_shipped_configuration_with_the_emulated_library_as_potential(golden_dir, tmp_path, deck):
    """examples/waterbox as shipped - NGLFCONSTRAINT, LANGEVIN groups and the barostat all run by ddcMD on the host - with the
    library as its MARTINI potential: the barostat's box reaches the library through ddcb200_setBox.  ras_small adds velocity
    constraints, which ddcMD applies through the residue table that martini() refreshes on every call: the binding keeps it."""
    check_shim(golden_dir, deck, tmp_path, SHIM_EMU, "full")


@pytest.mark.gpu
@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(SHIM)), reason="oracle/_ref binaries not built")
@pytest.mark.parametrize("deck,variant", [("waterbox", "full"), ("ras_small", "full")])
def test_ddcmd_integrators_with_the_library_as_potential(golden_dir, tmp_path, deck, variant):
    check_shim(golden_dir, deck, tmp_path, SHIM, variant)


@pytest.mark.gpu
@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(SHIM)), reason="oracle/_ref binaries not built")
@pytest.mark.parametrize("deck", ["popc_small", "ras_small", "waterbox"])
def test_ddcmd_with_the_library_plugged_in(golden_dir, tmp_path, deck):
    check_shim(golden_dir, deck, tmp_path, SHIM)
