"""Build / walk variants that must not change a result (collected after every other GPU test):

  * the one-pass cell list build against the two-pass build: bit-identical rows, forces, trajectories (auto mode picks by timing);
  * the per-bead list-walk bound against the global one: bitwise equal trajectories;
  * DDCB200_BIN_EDGES: another row order, same pair set, forces to rounding.
"""
import numpy as np
import pytest

import ddcmd_b200 as dd
from test_gpu_parity import DECKS, F_TOL, _force_err, _load, _pairkey

pytestmark = pytest.mark.gpu


def _rows_and_forces(golden_dir, name, mode, monkeypatch):
    monkeypatch.setenv("DDCB200_LISTBUILD", mode)
    sim, ref = _load(golden_dir, name)
    sim.ddcenergy(1)
    e = sim.energyInfo()
    st = sim.getState()
    pairs = sim.getPairs()          # decoded in row order: equal arrays = equal rows, entry for entry
    cells = sim.getCells()[0]
    sim.nglf(61)                    # across three rebuilds: the auto mode has timed both builds twice by then
    e2 = sim.energyInfo()
    st2 = sim.getState()
    info = sim.listBuildInfo()
    sim.close()
    return pairs, cells, st, e, st2, e2, info


@pytest.mark.parametrize("name", DECKS)
def test_list_builds_agree_bit_for_bit(golden_dir, name, monkeypatch):
    """The one-pass cell build (k_nbr_cell) and the two-pass build (k_nbr_filter + k_nbr_exact) write the same rows in the same
    order, so forces, energies and the trajectory across a rebuild are bitwise equal whichever one the timing picks."""
    a = _rows_and_forces(golden_dir, name, "twopass", monkeypatch)
    b = _rows_and_forces(golden_dir, name, "cell", monkeypatch)
    assert a[6][0] == 1 and b[6][0] == 2
    for x, y in zip(a[0], b[0]):
        assert np.array_equal(x, y)
    assert np.array_equal(a[1], b[1])
    for k in ("fx", "fy", "fz"):
        assert np.array_equal(a[2][k], b[2][k])
    assert a[3].eion == b[3].eion and a[3].nPairsListed == b[3].nPairsListed
    for k in ("rx", "ry", "rz", "vx", "vy", "vz", "fx"):
        assert np.array_equal(a[4][k], b[4][k])
    assert a[5].eion == b[5].eion and a[5].rk == b[5].rk and a[5].nPairsListed == b[5].nPairsListed
    # auto: the first four builds alternate, then the faster of the two
    c = _rows_and_forces(golden_dir, name, "auto", monkeypatch)
    assert c[6][0] in (1, 2) and c[6][1][0] > 0.0 and c[6][1][1] > 0.0
    assert np.array_equal(c[4]["rx"], a[4]["rx"]) and c[5].eion == a[5].eion


def test_bin_edges_knob_keeps_the_pair_set(golden_dir, monkeypatch):
    """DDCB200_BIN_EDGES only reorders the entries of a row (here: two bins instead of eight): same pairs, forces to rounding."""
    sim, ref = _load(golden_dir, "popc_small")
    sim.ddcenergy(1)
    a = sim.getState()
    pa = sim.getPairs()
    sim.close()
    monkeypatch.setenv("DDCB200_BIN_EDGES", "-0.25,-0.25,-0.25,0.25,0.25,0.25,0.25")
    sim, _ = _load(golden_dir, "popc_small")
    sim.ddcenergy(1)
    b = sim.getState()
    pb = sim.getPairs()
    assert _force_err(b, ref, "s0_") < F_TOL
    assert np.array_equal(np.sort(_pairkey(pa[0], pa[1])), np.sort(_pairkey(pb[0], pb[1])))
    assert not np.array_equal(pa[1], pb[1])              # the rows really are in a different order
    assert np.abs(a["fx"] - b["fx"]).max() <= 1e-10 * np.abs(a["fx"]).max()
    sim.nglf(25)                                         # displacement-bounded walk across a rebuild with the merged bins
    tr = ref["trace"].reshape(-1, 16)
    e = sim.energyInfo()
    etot = tr[24, 1] + tr[24, 2]
    assert abs((e.eion + e.rk) - etot) <= 1e-9 * max(abs(etot), abs(tr[24, 2]))
    sim.close()
    monkeypatch.setenv("DDCB200_BIN_EDGES", "0.5,0.1")
    with pytest.raises(dd.DdcError):
        _load(golden_dir, "popc_small")


@pytest.mark.parametrize("name", ["popc_small", "ras_small"])
def test_per_bead_walk_bound_is_bitwise_neutral(golden_dir, name, monkeypatch):
    """k_pair stops each row at rmax + dmax + the bead's own displacement instead of rmax + 2 dmax: the entries it no longer visits
    would have added exact zeros, so 45 steps (two rebuilds, growing displacements) are bitwise the same with either bound."""
    out = {}
    for mode in ("global", "bead"):
        monkeypatch.setenv("DDCB200_WALK", mode)
        sim, _ = _load(golden_dir, name)
        sim.nglf(45)
        e = sim.energyInfo()
        st = sim.getState()
        out[mode] = (st, e.eion, e.rk, np.array(e.virial[:]))
        sim.close()
    a, b = out["global"], out["bead"]
    for k in ("rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz"):
        assert np.array_equal(a[0][k], b[0][k]), k
    assert a[1] == b[1] and a[2] == b[2] and np.array_equal(a[3], b[3])


@pytest.mark.parametrize("mode", ["twopass", "cell"])
def test_row_capacity_regrow(golden_dir, monkeypatch, mode):
    """A first build whose rows overflow the allocated capacity (forced small here) regrows from the measured maximum and repeats:
    same pairs, same forces as with the default capacity, in both builds."""
    monkeypatch.setenv("DDCB200_LISTBUILD", mode)
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


def test_auto_mode_self_check_falls_back(golden_dir, monkeypatch, capfd):
    """auto mode compares its first one-pass build with a two-pass build of the same state; a (here: injected) difference makes
    the context keep the two-pass build, with the same results."""
    monkeypatch.setenv("DDCB200_LISTBUILD", "twopass")
    sim, ref = _load(golden_dir, "popc_small")
    sim.nglf(45)
    a = sim.getState()
    ea = sim.energyInfo()
    sim.close()
    monkeypatch.setenv("DDCB200_LISTBUILD", "auto")
    monkeypatch.setenv("DDCB200_SELFCHECK_FAULT", "1")
    sim, _ = _load(golden_dir, "popc_small")
    sim.nglf(45)
    b = sim.getState()
    eb = sim.energyInfo()
    assert sim.listBuildInfo()[0] == 1
    sim.close()
    assert "keeping the two-pass build" in capfd.readouterr().err
    for k in ("rx", "vx", "fx", "fz"):
        assert np.array_equal(a[k], b[k])
    assert ea.eion == eb.eion and ea.nPairsListed == eb.nPairsListed
    # without the injected fault the check passes silently and both builds get timed
    monkeypatch.delenv("DDCB200_SELFCHECK_FAULT")
    sim, _ = _load(golden_dir, "popc_small")
    sim.nglf(65)
    info = sim.listBuildInfo()
    c = sim.getState()
    sim.close()
    assert info[0] in (1, 2) and info[1][0] > 0 and info[1][1] > 0
    assert "keeping the two-pass build" not in capfd.readouterr().err
    assert np.array_equal(c["rx"][:10], c["rx"][:10])
