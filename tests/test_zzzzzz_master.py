"""simulateMaster through the C-ABI: the whole run of a deck - firstEnergyCall, the MD loop with
the reference's print / checkpoint cadence, the `data` file, ddcMD_CMDS control and ddcMD-format restarts - against what the
UNMODIFIED reference wrote for the same deck (tests/golden/snapshot.json, snapshot_run.npz; made by
tests/golden/make_snapshot_golden.py).  Collected after the step parity tests."""
import json
import os
import re
import shutil

import numpy as np
import pytest

import ddcmd_b200 as dd
import nglfc_decks

pytestmark = pytest.mark.gpu


def stage_run(golden_dir, deck, variant, tmp_path, deltaloop=10):
    if variant:
        d = nglfc_decks.make_variant(golden_dir, deck, variant, tmp_path)
    else:
        d = os.path.join(str(tmp_path), deck)
        shutil.copytree(os.path.join(golden_dir, deck), d, symlinks=True)
    p = os.path.join(d, "object.data")
    s = open(p).read()
    s = re.sub(r"deltaloop=\d+;", "deltaloop=%d;" % deltaloop, s)
    s = re.sub(r"printrate=\d+;", "printrate=5;", s)
    s = re.sub(r"checkpointrate=\d+;", "checkpointrate=10;", s)
    open(p, "w").write(s)
    return d


def read_snapshot(snap):
    raw = open(os.path.join(snap, "atoms#000000"), "rb").read()
    header = raw[:raw.index(b"}")].decode()
    lrec = int(re.search(r"lrec=(\d+)", header).group(1))
    recs = raw[raw.index(b"\n\n", raw.index(b"}")) + 2:]
    n = len(recs) // lrec
    gid = np.array([int(recs[i * lrec:(i + 1) * lrec].split()[1], 16) for i in range(n)], np.uint64)
    rv = np.array([[float(x) for x in recs[i * lrec:(i + 1) * lrec].split()[5:11]] for i in range(n)])
    return header, gid, rv


def check_master(golden_dir, deck, variant, tmp_path):
    key = deck + ("_" + variant if variant else "")
    g = json.load(open(os.path.join(golden_dir, "snapshot.json")))[key]["run"]
    d = stage_run(golden_dir, deck, variant, tmp_path)
    dd.simulateMaster(os.path.join(d, "object.data"))
    # ---- the data file: same header, same loops, numbers to the step-parity tolerances
    ours, ref = open(os.path.join(d, "data")).read().splitlines(), g["data"].splitlines()
    assert ours[0] == ref[0]
    assert len(ours) == len(ref) == 4
    cons = variant is not None and deck == "ras_small"
    for a, b in zip(ours[1:], ref[1:]):
        fa, fb = a.split(), b.split()
        assert fa[0] == fb[0] and len(fa) == len(fb) == 11
        va, vb = np.array(fa[1:], float), np.array(fb[1:], float)
        etol = 1e-7 if cons else 1e-9
        scale = max(abs(vb[1]), abs(vb[2]))
        assert abs(va[0] - vb[0]) < 1e-9                       # time
        assert np.all(np.abs(va[1:4] - vb[1:4]) <= etol * scale + 2e-12)     # Etotal, Ekin, Epot (12 decimals printed)
        assert abs(va[4] - vb[4]) <= etol * abs(vb[4]) + 2e-8   # temperature
        assert abs(va[5] - vb[5]) <= 1e-6 * max(abs(vb[5]), 100.0)    # pressure: a difference of large virial terms
        assert np.all(np.abs(va[6:] - vb[6:]) <= 1e-9 * np.abs(vb[6:]) + 2e-8)   # volume per bead, box edges
    # ---- the checkpoint at loop 10
    snap = os.path.join(d, "snapshot.000000000010")
    assert os.readlink(os.path.join(d, "restart")) == g["restart_link"]
    strip = lambda t: re.sub(r"run_id=0x[0-9a-f]{8}", "run_id=X", t)          # noqa: E731
    tr, to = strip(g["restart"]).split(), strip(open(os.path.join(snap, "restart")).read()).split()
    assert len(tr) == len(to)
    for x, y in zip(to, tr):
        try:
            fx, fy = float(x.rstrip(";")), float(y.rstrip(";"))
        except ValueError:
            assert x == y
        else:
            assert abs(fx - fy) <= 1e-11 * abs(fy)
    header, gid, rv = read_snapshot(snap)
    assert "loop=10;" in header and "lrec=%d;" % g["lrec"] in header
    if deck == "popc_small":
        z = np.load(os.path.join(golden_dir, "snapshot_run.npz"))
        assert np.array_equal(gid, z[key + "_gid"])
        ref_rv = z[key + "_rv"]
        h = np.array([float(x) for x in ref[-1].split()[8:11]])
        dr = rv[:, :3] - ref_rv[:, :3]
        dr -= h * np.rint(dr / h)
        assert np.abs(dr).max() < 1e-8
        assert np.abs(rv[:, 3:] - ref_rv[:, 3:]).max() <= 1e-8 * np.abs(ref_rv[:, 3:]).max()
    # ---- the restart written here starts the next run (loop 10 -> 15) and a ddcMD_CMDS "exit" stops and checkpoints it
    p = os.path.join(d, "object.data")
    s = open(p).read()
    open(p, "w").write(re.sub(r"deltaloop=\d+;", "deltaloop=1000;", s))
    open(os.path.join(d, "ddcMD_CMDS"), "w").write("exit\n")
    dd.simulateMaster(p)
    lines = open(os.path.join(d, "data")).read().splitlines()
    assert [ln.split()[0] for ln in lines[-2:]] == ["000000000010", "000000000015"]
    assert os.path.getsize(os.path.join(d, "ddcMD_CMDS")) == 0
    assert os.readlink(os.path.join(d, "restart")) == "./snapshot.000000000015/restart"
    d2 = dd.Deck(p)
    assert int(d2.s.loop) == 15
    # the restarted trajectory continues the first one: loop-10 line re-printed from the restart file agrees to print precision of the file
    a, b = np.array(lines[-2].split()[1:], float), np.array(ours[-1].split()[1:], float)
    assert np.all(np.abs(a[1:4] - b[1:4]) <= 1e-9 * max(abs(b[1]), abs(b[2])))


# ("waterbox", "full") is the configuration the reference ships in examples/waterbox (NGLFCONSTRAINT + LANGEVIN groups + barostat)
@pytest.mark.parametrize("deck,variant", [("popc_small", None), ("popc_small", "full"), ("ras_small", "full"), ("waterbox", None), ("waterbox", "full")])
def test_simulateMaster_matches_reference_run(golden_dir, tmp_path, deck, variant):
    check_master(golden_dir, deck, variant, tmp_path)


def check_subset(golden_dir, tmp_path):
    """ANALYSIS type = subsetWrite, format = binaryCharmm: the files the reference wrote at loops 5 and 10, record for record."""
    mg = _golden_module(golden_dir)
    g = json.load(open(os.path.join(golden_dir, "snapshot.json")))["popc_small_subset"]
    z = np.load(os.path.join(golden_dir, "snapshot_run.npz"))
    d = os.path.join(str(tmp_path), "popc_small")
    shutil.copytree(os.path.join(golden_dir, "popc_small"), d, symlinks=True)
    p = os.path.join(d, "object.data")
    text = open(p).read()
    open(p, "w").write(mg.subset_deck(text))
    deck = dd.Deck(p)
    assert int(deck.s.nSubsets) == 1
    dd.simulateMaster(p)
    for loop in (5, 10):
        raw = open(os.path.join(d, "snapshot.%012d" % loop, "pos#000000"), "rb").read()
        k = raw.index(b"}")
        rec = np.frombuffer(raw[raw.index(b"\n\n", k) + 2:], dtype=mg.SUBSET_DTYPE)
        assert np.array_equal(rec["gid"], z["subset_%d_gid" % loop]) and np.array_equal(rec["pin"], z["subset_%d_pin" % loop])
        # float32 of positions that agree to ~1e-12: equal, or one float ulp apart where the double sits on a rounding boundary
        assert np.abs(rec["r"] - z["subset_%d_r" % loop]).max() <= 1e-5
        assert (rec["r"] != z["subset_%d_r" % loop]).mean() < 0.01
        strip = lambda t: [re.sub(r"create_time=[^;]*;|run_id=0x[0-9a-f]{8};", "", x) for x in t.splitlines() if not x.startswith("code_version")]   # noqa: E731
        assert strip(raw[:k].decode()) == strip(g["header_%d" % loop])
    # unsupported analyses are refused at load time, loudly
    open(p, "w").write(mg.subset_deck(text).replace("format=binaryCharmm;", "format=ovito;"))
    with pytest.raises(dd.DdcError, match="subsetWrite with format = binaryCharmm, is supported"):
        dd.Deck(p)


def test_subsetWrite_matches_reference(golden_dir, tmp_path):
    check_subset(golden_dir, tmp_path)


def _golden_module(golden_dir):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_snapshot_golden", os.path.join(golden_dir, "make_snapshot_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg


def check_paircorr(golden_dir, tmp_path):
    """ANALYSIS type = PAIRCORRELATION: (1) the g(r) table the reference wrote after sampling loops 5 and 10 of popc_small;
    (2) the device counts at loop 0 against a brute-force count over all pairs, linear and logarithmic bins."""
    mg = _golden_module(golden_dir)
    g = json.load(open(os.path.join(golden_dir, "snapshot.json")))["popc_small_gr"]
    table = np.load(os.path.join(golden_dir, "snapshot_run.npz"))["gr_table"]
    d = os.path.join(str(tmp_path), "popc_small")
    shutil.copytree(os.path.join(golden_dir, "popc_small"), d, symlinks=True)
    p = os.path.join(d, "object.data")
    text = open(p).read()
    open(p, "w").write(mg.paircorr_deck(text))
    dd.simulateMaster(p)
    lines = open(os.path.join(d, "snapshot.%012d" % 10, "gr.dat")).read().splitlines()
    assert lines[:3] == g["header"]
    got = np.array([[float(x) for x in ln.split()] for ln in lines[3:]])
    assert got.shape == table.shape
    # positions agree with the reference's to ~1e-12, so a pair sitting on a bin edge may fall on the other side: all but a few of
    # the 22 x 406 entries are identical as printed
    assert np.array_equal(got[:, 0], table[:, 0])
    assert (got != table).sum() <= 6
    assert np.abs(got - table).max() <= 0.05 * max(table[:, 1:].max(), 1.0)
    assert not os.path.exists(os.path.join(d, "snapshot.%012d" % 5, "gr.dat"))
    # ---- direct call against brute force
    sim = dd.simulate_init(os.path.join(golden_dir, "popc_small", "object.data"))
    sim.ddcenergy(1)
    sim.nglf(7)                                        # displaced since the build: the widened cell walk must still find every pair
    st = sim.getState()
    r = np.stack([st["rx"], st["ry"], st["rz"]], 1)
    h = np.array([sim.getBox()[k] for k in (0, 4, 8)])
    sp = sim.deck.array("species").astype(np.int64)
    gid = sim.deck.array("gid")
    ns = int(sim.deck.s.nspecies)
    A = dd.units_convert(1.0, "Angstrom", None)
    n = len(r)
    ii, jj = np.triu_indices(n, 1)
    dr = r[ii] - r[jj]
    dr -= h * np.rint(dr / h)
    dist = np.sqrt((dr ** 2).sum(1))

    def combo(a, b):
        mx, mn = np.maximum(a, b), np.minimum(a, b)
        return (mx - mn) + ns * mn - (mn * (mn - 1)) // 2

    for log_scale, rmin, nb, rmax in ((False, 0.0, 26, 13.0 * A), (True, 3.0 * A, 12, 14.0 * A)):
        delta = (np.log10(rmax) - np.log10(rmin)) / nb if log_scale else (rmax - rmin) / nb
        counts, natoms = sim.pairCorrelation(nb, rmin, delta, rmax, log_scale)
        assert np.array_equal(natoms, np.bincount(sp, minlength=ns).astype(np.uint64))
        q = (np.log10(dist) - np.log10(rmin)) / delta if log_scale else (dist - rmin) / delta
        ok = (dist < rmax) & (q >= 0) & (q < nb)
        want = np.zeros((ns * (ns + 1) // 2, nb), np.int64)
        np.add.at(want, (combo(sp[ii[ok]], sp[jj[ok]]), q[ok].astype(np.int64)), np.where(sp[ii[ok]] == sp[jj[ok]], 2, 1))
        diff = np.abs(counts.astype(np.int64) - want)
        assert counts.sum() > 100000 and diff.sum() <= 4, (counts.sum(), diff.sum())     # <= 2 pairs on a bin edge may round the other way
    assert len(np.unique(gid)) == n
    sim.close()


def test_paircorrelation_matches_reference(golden_dir, tmp_path):
    check_paircorr(golden_dir, tmp_path)


def check_kinetic_classes(golden_dir, tmp_path):
    """per-GROUP / per-SPECIES kinetic terms and thermal flux (kinetic_terms, src/energy.c:116-143) against numpy on the same state"""
    d = nglfc_decks.make_variant(golden_dir, "popc_small", "lang", tmp_path)      # two GROUP objects (every bead names the first), 28 species
    sim = dd.simulate_init(os.path.join(d, "object.data"))
    sim.ddcenergy(1)
    sim.eval_integrator(7)
    e = sim.energyInfo()
    st = sim.getState()
    dk = sim.deck
    sp = dk.array("species").astype(np.int64)
    m = dk.array("specMass")[sp]
    grp = dk.array("groupOfBead").astype(np.int64)
    v = np.stack([st["vx"], st["vy"], st["vz"]], 1)
    K = 0.5 * m * (v ** 2).sum(1)
    for by_species, cls, n in ((False, grp, int(dk.s.nGroups)), (True, sp, int(dk.s.nspecies))):
        got = sim.kineticByClass(by_species)
        assert got.shape == (n, 12)
        want = np.zeros((n, 12))
        np.add.at(want[:, 0], cls, K)
        np.add.at(want[:, 1], cls, m)
        np.add.at(want[:, 2], cls, 1.0)
        for k, (a, b) in enumerate(((0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2))):
            np.add.at(want[:, 3 + k], cls, m * v[:, a] * v[:, b])
        for a in range(3):
            np.add.at(want[:, 9 + a], cls, K * v[:, a])
        scale = np.abs(want).max(0) + 1e-300
        assert np.all(np.abs(got - want) <= 1e-12 * scale)
        assert np.array_equal(got[:, 2], want[:, 2])
        assert abs(got[:, 0].sum() - e.rk) <= 1e-12 * e.rk                        # the classes partition the system
        assert np.allclose(got[:, 3:9].sum(0), np.array(e.tion[:]), rtol=1e-12, atol=1e-12 * abs(e.tion[0]))
    sim.close()


def test_kinetic_terms_by_group_and_species(golden_dir, tmp_path):
    check_kinetic_classes(golden_dir, tmp_path)
