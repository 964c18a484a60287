"""Walk / row-order variants that must not change a result:

  * the per-bead list-walk bound against the global one: bitwise equal trajectories;
  * DDCB200_NEAR: another row order, same pair set, forces to rounding;
  * DDCB200_PRUNE: the pruned rows of the pair walk, bitwise the results of the full walk, in every integrator / rebuild mode;
  * a first build whose rows overflow the allocated capacity regrows and repeats.
"""
import os

import numpy as np
import pytest

import ddcmd_b200 as dd
from test_gpu_parity import F_TOL, _force_err, _load, _pairkey

pytestmark = pytest.mark.gpu


def test_near_edge_knob_keeps_the_pair_set(golden_dir, monkeypatch):
    """DDCB200_NEAR moves the edge between the two segments of a row: same pairs, another order inside the rows, forces to rounding."""
    sim, ref = _load(golden_dir, "popc_small")
    sim.ddcenergy(1)
    a = sim.getState()
    pa = sim.getPairs()
    sim.close()
    monkeypatch.setenv("DDCB200_NEAR", "0.6")
    sim, _ = _load(golden_dir, "popc_small")
    sim.ddcenergy(1)
    b = sim.getState()
    pb = sim.getPairs()
    assert _force_err(b, ref, "s0_") < F_TOL
    assert np.array_equal(np.sort(_pairkey(pa[0], pa[1])), np.sort(_pairkey(pb[0], pb[1])))
    assert not np.array_equal(pa[1], pb[1])              # the rows really are in a different order
    assert np.abs(a["fx"] - b["fx"]).max() <= 1e-10 * np.abs(a["fx"]).max()
    sim.nglf(25)                                         # across a rebuild and several prunes
    tr = ref["trace"].reshape(-1, 16)
    e = sim.energyInfo()
    etot = tr[24, 1] + tr[24, 2]
    assert abs((e.eion + e.rk) - etot) <= 1e-9 * max(abs(etot), abs(tr[24, 2]))
    sim.close()
    monkeypatch.setenv("DDCB200_NEAR", "1.5")
    with pytest.raises(dd.DdcError):
        _load(golden_dir, "popc_small")


@pytest.mark.parametrize("name", ["popc_small", "ras_small"])
def test_per_bead_walk_bound_is_bitwise_neutral(golden_dir, name, monkeypatch):
    """k_pair stops each row at rmax + the bead's own displacement + the largest displacement in its stencil cells (or, "bead", of
    any resident bead) instead of rmax + 2 dmax: the entries it no longer visits would have added exact zeros, so 45 steps (two
    rebuilds, growing displacements) are bitwise the same with every bound."""
    out = {}
    for mode in ("global", "bead", "cell"):
        monkeypatch.setenv("DDCB200_WALK", mode)
        sim, _ = _load(golden_dir, name)
        sim.nglf(45)
        e = sim.energyInfo()
        st = sim.getState()
        out[mode] = (st, e.eion, e.rk, np.array(e.virial[:]))
        sim.close()
    a = out["global"]
    for b in (out["bead"], out["cell"]):
        for k in ("rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz"):
            assert np.array_equal(a[0][k], b[0][k]), k
        assert a[1] == b[1] and a[2] == b[2] and np.array_equal(a[3], b[3])


@pytest.mark.parametrize("name", ["popc_small", "ras_small"])
def test_pruned_rows_are_bitwise_neutral(golden_dir, name, monkeypatch):
    """DDCB200_PRUNE: every few steps the pair walk also writes a shorter row per bead (entries closer than rmax + margin) and the
    steps in between walk that row while the displacement bounds since the prune stay within the margin, the full row otherwise.
    Every skipped entry is outside the cutoff, so 45 steps (two rebuilds) are bitwise the same - with a margin that holds, with
    one so small that beads fall back to their full rows, and with energies evaluated on prune steps and on steps in between."""
    out = {}
    for tag, prune in (("off", "0"), ("p4", "4"), ("p5wide", "5,0.6"), ("p3tiny", "3,0.01"), ("p2", "2,0.1")):
        monkeypatch.setenv("DDCB200_PRUNE", prune)
        sim, _ = _load(golden_dir, name)
        trace = []
        for n in (7, 6, 12, 20):                  # energies at steps 7, 13, 25, 45: prune steps and steps in between
            sim.nglf(n)
            e = sim.energyInfo()
            trace.append((e.eion, e.rk, tuple(e.virial[:])))
        st = sim.getState()
        info = sim.pruneInfo()
        out[tag] = (st, trace, info)
        sim.close()
    # the pruned rows really are in use: with a margin that holds nearly every bead would walk its short row, with the tiny one none
    i4, itiny = out["p4"][2], out["p3tiny"][2]
    n = len(out["off"][0]["rx"])
    assert out["off"][2]["every"] == 0 and out["off"][2]["since"] == -1
    assert i4["every"] == 4 and 0 <= i4["since"] < 4 and i4["full_entries"] > 0
    assert 0 < i4["pruned_entries"] < 0.8 * i4["full_entries"]
    assert i4["beads_pruned"] > 0.5 * n and i4["walk_next"] < 0.9 * i4["full_entries"]
    assert itiny["beads_pruned"] < 0.5 * n
    a = out["off"]
    for tag in ("p4", "p5wide", "p3tiny", "p2"):
        b = out[tag]
        for k in ("rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz"):
            assert np.array_equal(a[0][k], b[0][k]), (tag, k)
        assert a[1] == b[1], tag
    monkeypatch.setenv("DDCB200_PRUNE", "-1")
    with pytest.raises(dd.DdcError):
        _load(golden_dir, name)


@pytest.mark.parametrize("mode", ["barostat", "ur0"])
def test_pruned_rows_with_barostat_and_displacement_rebuilds(golden_dir, mode, tmp_path, monkeypatch):
    """The pruned rows also serve NGLFCONSTRAINT with the barostat (the box the displacement bounds refer to is the one of the last
    prune) and DDC updateRate = 0 (neighborCheck keeps the positions of the build, the prunes move their own reference on):
    30 steps bitwise equal with and without them, same rebuild loops."""
    import nglfc_decks
    from test_zzzz_ur0 import ur0_deck
    out = {}
    for prune in ("0", "4", "3,0.05"):
        monkeypatch.setenv("DDCB200_PRUNE", prune)
        if mode == "barostat":
            d = nglfc_decks.make_variant(golden_dir, "popc_small", "full", tmp_path / ("b" + prune.replace(",", "_")))
            sim = dd.simulate_init(os.path.join(d, "object.data"))
        else:
            sim = dd.simulate_init(ur0_deck(golden_dir, "popc_small", tmp_path / ("u" + prune.replace(",", "_"))))
        sim.ddcenergy(1)
        builds, es = [], []
        for _ in range(30):
            sim.eval_integrator(1)
            e = sim.energyInfo()
            builds.append(sim.lastListBuild())
            es.append((e.eion, e.rk, tuple(e.virial[:])))
        out[prune] = (sim.getState(), builds, es, tuple(sim.getBox()), sim.pruneInfo())
        sim.close()
    a = out["0"]
    assert out["4"][4]["every"] == 4 and out["4"][4]["pruned_entries"] > 0
    for tag in ("4", "3,0.05"):
        b = out[tag]
        assert a[1] == b[1] and a[2] == b[2] and a[3] == b[3], tag
        for k in ("rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz"):
            assert np.array_equal(a[0][k], b[0][k]), (tag, k)


def test_row_capacity_regrow(golden_dir, monkeypatch):
    """A first build whose rows overflow the allocated capacity (forced small here) regrows from the measured maximum and repeats:
    same pairs, same forces as with the default capacity."""
    sim, ref = _load(golden_dir, "popc_small")
    sim.ddcenergy(1)
    a = sim.getState()
    pa = sim.getPairs()
    sim.close()
    monkeypatch.setenv("DDCB200_NBRCAP", "40")
    sim, _ = _load(golden_dir, "popc_small")
    sim.ddcenergy(1)
    b = sim.getState()
    pb = sim.getPairs()
    assert len(pb[0]) == int(ref["npairs"][0])
    for x, y in zip(pa, pb):
        assert np.array_equal(x, y)
    for k in ("fx", "fy", "fz"):
        assert np.array_equal(a[k], b[k])
    sim.nglf(21)
    assert sim.energyInfo().nPairsListed == int(ref["trace"].reshape(-1, 16)[20, 14])
    sim.close()
