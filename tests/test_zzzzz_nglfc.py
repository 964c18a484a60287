"""NGLFCONSTRAINT integrator: LANGEVIN groups with the per-bead LCG64 streams, velocity
constraints and the Berendsen molecular-pressure barostat against the reference's nglfconstraint
(src/nglfconstraint.c:510-574).  Golden: tests/golden/nglfc.npz (tests/golden/make_nglfc_golden.py); `waterbox full` is
the configuration the reference ships in examples/waterbox/object.data.  Collected after the NGLF parity tests."""
import os

import numpy as np
import pytest

import ddcmd_b200 as dd
import nglfc_decks

pytestmark = pytest.mark.gpu

CASES = [("waterbox", "full"), ("popc_small", "lang"), ("popc_small", "baro"), ("popc_small", "full"), ("ras_small", "lang"),
         ("ras_small", "full")]


def check_nglfc(golden_dir, deck, variant, tmp_path, nsteps=40):
    g = np.load(os.path.join(golden_dir, "nglfc.npz"))
    key = "%s_%s_" % (deck, variant)
    tr, box = g[key + "trace"], g[key + "box"]
    lang, baro = nglfc_decks.VARIANTS[variant]
    d = nglfc_decks.make_variant(golden_dir, deck, variant, tmp_path)
    sim = dd.simulate_init(os.path.join(d, "object.data"))
    assert int(sim.deck.s.integratorType) == 1
    # velocity constraints grow the rounding differences of the two summation orders faster than plain dynamics does
    cons = int(sim.deck.s.nCons) > 0
    etol, xtol = (1e-9, 1e-7) if cons else (1e-10, 1e-9)
    sim.ddcenergy(1)
    for s in range(nsteps):
        sim.eval_integrator(1)
        e = sim.energyInfo()
        assert e.loop == int(tr[s, 0])
        assert e.nPairsListed == int(tr[s, 14]), "pairs listed at loop %d" % e.loop
        etot = tr[s, 1] + tr[s, 2]
        assert abs((e.eion + e.rk) - etot) <= etol * max(abs(etot), abs(tr[s, 2])), "Etot at loop %d" % e.loop
        assert abs(e.rk - tr[s, 2]) <= etol * abs(tr[s, 2]), "kinetic energy at loop %d" % e.loop
        h = sim.getBox()
        assert np.allclose([h[0], h[4], h[8]], box[s], rtol=1e-12, atol=0), "box at loop %d" % e.loop   # changeVolume
        if not baro:
            assert h[0] == box[s, 0] and h[8] == box[s, 2]
    if nsteps == len(tr):
        st = sim.getState()
        assert np.abs(st["rz"] - g[key + "rz"]).max() <= xtol * np.abs(g[key + "rz"]).max()
        assert np.abs(st["vz"] - g[key + "vz"]).max() <= xtol * 100 * np.abs(g[key + "vz"]).max()
        if lang:
            assert np.array_equal(sim.getRandom(), g[key + "rng"])     # every bead drew exactly the reference's numbers
    assert sim.constraintFailures() == 0
    sim.close()


@pytest.mark.parametrize("deck,variant", CASES)
def test_nglfconstraint_matches_reference(golden_dir, deck, variant, tmp_path):
    check_nglfc(golden_dir, deck, variant, tmp_path)


def test_nglf_with_langevin_groups_is_the_unconstrained_pass(golden_dir, tmp_path):
    """INTEGRATOR NGLF + LANGEVIN groups (nglf calls group->velocityUpdate, src/nglf.c:75,104) = NGLFCONSTRAINT with beta = 0
    when the deck has no constraints: same golden as popc_small lang."""
    g = np.load(os.path.join(golden_dir, "nglfc.npz"))
    tr = g["popc_small_lang_trace"]
    d = nglfc_decks.make_variant(golden_dir, "popc_small", "lang", tmp_path)
    p = os.path.join(d, "object.data")
    s = open(p).read()
    i0 = s.index("nglf INTEGRATOR")
    i1 = s.index("\n", i0)
    open(p, "w").write(s[:i0] + "nglf INTEGRATOR { type = NGLF; }" + s[i1:])
    sim = dd.simulate_init(p)
    assert int(sim.deck.s.integratorType) == 0
    sim.ddcenergy(1)
    sim.eval_integrator(10)
    e = sim.energyInfo()
    etot = tr[9, 1] + tr[9, 2]
    assert abs((e.eion + e.rk) - etot) <= 1e-10 * max(abs(etot), abs(tr[9, 2]))
    sim.close()


@pytest.mark.parametrize("nproc,deck,variant,lattice", [(2, "ras_small", "full", ()), (2, "waterbox", "full", ()), (4, "popc_small", "full", (2, 2, 1))])
def test_nglfconstraint_on_several_gpus(nproc, deck, variant, lattice):
    """The same traces with the system decomposed over 2 / 4 GPUs (tests/mgpu_nglfc_worker.py)."""
    import subprocess
    import sys
    if dd.lib().ddcb200_deviceCount() < nproc:
        pytest.skip("needs %d GPUs" % nproc)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % nproc, "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(root, "tests", "mgpu_nglfc_worker.py"), deck, variant] + [str(x) for x in lattice]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MGPU_NGLFC_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-6000:]
