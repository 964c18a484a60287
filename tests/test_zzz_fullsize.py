"""Parity and size-independent properties at BASELINE.json's configurations.

  popc_100k / ras_140k   the single-GPU configs: the UNMODIFIED reference CPU path (oracle/_ref/ref_dump, built by
                         oracle/build_ref.sh) is run live on the same generated deck - it finishes in seconds at this size -
                         and the CUDA path must meet the same bars as on the small golden decks: cells and pair count
                         bit-exact, forces 1e-6, energies 1e-9, and the same energies after two integration steps.
  membrane_1m            the bench workload: the same comparison with the live reference (the two pair lists through an
                         order-independent hash of their (gid, gid) pairs instead of 400 MB of indices), plus properties that hold
                         at any size - Newton's third law over the full (both-direction) lists (sum of pair forces = 0), bitwise
                         run-to-run reproducibility, and energy conservation over a rebuild.

The decks come from ddcmd_b200.synth (fixed seeds) and are cached under DDCB200_DECK_CACHE like bench.py's.
Collected after the golden-deck parity tests."""
import os
import subprocess
import sys

import numpy as np
import pytest

import ddcmd_b200 as dd
from ddcmd_b200 import synth
from refdump import read_records

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DUMP = os.path.join(ROOT, "oracle", "_ref", "ref_dump")
CACHE = os.environ.get("DDCB200_DECK_CACHE", "/tmp/ddcb200_decks")


def get_deck(name):
    path = os.path.join(CACHE, name)
    if not os.path.exists(os.path.join(path, "snapshot.mem", "atoms#000000")):
        synth.make(name).write_deck(path)
    return path


def run_oracle(path, nsteps, mode=""):
    out = os.path.join(path, "_fullsize.bin")
    subprocess.check_call(["bash", "-c", "ulimit -s unlimited; exec '%s' '%s' %d 0 %s" % (REF_DUMP, out, nsteps, mode)], cwd=path,
                          stdout=open(os.path.join(path, "_fullsize.log"), "w"), stderr=subprocess.STDOUT)
    r = read_records(out)
    os.remove(out)
    return r


def pair_key(a, b):
    a = a.astype(np.int64)
    b = b.astype(np.int64)
    return (np.minimum(a, b) << 32) | np.maximum(a, b)


def check_against_live_oracle(name, nsteps=2, hashed=False):
    """hashed: the reference dumps order-independent hashes of its two pair lists instead of the lists (million-bead decks);
    otherwise the full pair SETS are compared."""
    path = get_deck(name)
    ref = run_oracle(path, nsteps, "hash" if hashed else "")
    sim = dd.simulate_init(os.path.join(path, "object.data"))
    n = sim.deck.n
    assert n == int(ref["nion"][0])
    sim.ddcenergy(1)
    # cells and list: bit-exact
    cell, dims, geom = sim.getCells()
    assert list(dims) == list(ref["geom_dims"][:3])
    assert np.array_equal(geom, ref["geom_parms"][:9])
    assert np.array_equal(cell, ref["cell"])
    e = sim.energyInfo()
    assert e.nPairsListed == int(ref["npairs"][0])
    if hashed:
        # count, sum and xor (mod 2^64) of a 64-bit mix of every (gid, gid) pair, interacting list then pruned list
        assert np.array_equal(sim.pairSetHash(), ref["pairhash"].view(np.uint64)), (sim.pairSetHash(), ref["pairhash"])
    else:
        bi, bj, pr = sim.getPairs()
        p0, p1 = ref["pairs0"].reshape(-1, 2), ref["pairs1"].reshape(-1, 2)
        assert np.array_equal(np.sort(pair_key(bi[pr == 0], bj[pr == 0])), np.sort(pair_key(p0[:, 0], p0[:, 1])))
        assert np.array_equal(np.sort(pair_key(bi[pr == 1], bj[pr == 1])), np.sort(pair_key(p1[:, 0], p1[:, 1])))
        # and the device-side hash agrees with the same hash of the reference's lists
        lab = ref["s0_label"].view(np.uint64)
        want = np.zeros(6, np.uint64)
        for l, pp in enumerate((p0, p1)):
            if len(pp):
                v = pair_hash(lab[pp[:, 0]], lab[pp[:, 1]])
                want[3 * l:3 * l + 3] = (len(pp), v.sum(dtype=np.uint64), np.bitwise_xor.reduce(v))
        assert np.array_equal(sim.pairSetHash(), want)
    # forces 1e-6, energies 1e-9
    st = sim.getState()
    f = np.stack([st["fx"], st["fy"], st["fz"]], 1)
    fr = np.stack([ref["s0_fx"], ref["s0_fy"], ref["s0_fz"]], 1)
    rms = np.sqrt((fr ** 2).sum(1).mean())
    err = (np.sqrt(((f - fr) ** 2).sum(1)) / np.maximum(np.sqrt((fr ** 2).sum(1)), rms)).max()
    assert err < 1e-6, err
    en = ref["s0_energy"]
    escale = max(abs(en[0]), 1e-3 * np.abs(ref["s0_fx"]).sum())
    assert abs(e.eion - en[0]) <= 1e-9 * escale
    assert np.allclose(np.array(e.virial[:]), en[6:12], rtol=1e-9, atol=1e-9 * np.abs(en[6:9]).max())
    be = ref["bioEnergies"]
    for got, want in ((e.eBond, be[0]), (e.eAngle, be[1]), (e.eTorsion, be[3]), (e.eImproper, be[4])):
        assert np.isclose(got, want, rtol=1e-9, atol=1e-12)
    # two integration steps
    tr = ref["trace"].reshape(-1, 16)
    sim.nglf(nsteps)
    e2 = sim.energyInfo()
    etot = tr[nsteps - 1, 1] + tr[nsteps - 1, 2]
    assert abs((e2.eion + e2.rk) - etot) <= 1e-9 * max(abs(etot), abs(tr[nsteps - 1, 2]))
    assert abs(e2.rk - tr[nsteps - 1, 2]) <= 1e-9 * abs(tr[nsteps - 1, 2])
    st2 = sim.getState()
    assert np.abs(st2["rx"] - ref["sN_rx"]).max() < 1e-10 and np.abs(st2["vz"] - ref["sN_vz"]).max() < 1e-13
    sim.close()
    return n


def mix64(x):
    x = x.astype(np.uint64)
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(30); x *= np.uint64(0xbf58476d1ce4e5b9)
        x ^= x >> np.uint64(27); x *= np.uint64(0x94d049bb133111eb)
        x ^= x >> np.uint64(31)
    return x


def pair_hash(ga, gb):
    lo, hi = np.minimum(ga, gb), np.maximum(ga, gb)
    with np.errstate(over="ignore"):
        return mix64(mix64(lo) + np.uint64(0x9e3779b97f4a7c15) * hi)


def check_properties(name, nsteps=25):
    path = get_deck(name)
    deck = dd.Deck(os.path.join(path, "object.data"))
    sim = dd.Simulate(deck)
    # non-bonded + bonded forces are internal forces: they sum to zero (restraints, absent from these decks, would not).
    # The pair kernel evaluates every pair from both ends independently, so this checks the full list is symmetric.
    sim.ddcenergy(1)
    st = sim.getState()
    e0 = sim.energyInfo()
    for k in ("fx", "fy", "fz"):
        assert abs(st[k].sum()) <= 1e-11 * np.abs(st[k]).sum(), k
    assert np.all(np.isfinite(st["fx"])) and np.isfinite(e0.eion)
    # the listed pair count is the half list; the virial is symmetric by construction: compare its trace with -sum r.f over
    # nearest images is not size-independent, so use the pressure identity instead: P V = (2 KE + W)/3, W = trace(virial)
    # (eval_energyInfo, src/energyInfo.c:75-148)
    vir = np.array(e0.virial[:])
    tion = np.array(e0.tion[:])
    assert abs(e0.pion - (vir[:3].sum() + tion[:3].sum()) / (3.0 * e0.volume)) <= 1e-12 * abs(e0.pion) + 1e-18
    # bitwise reproducibility across contexts
    sim2 = dd.Simulate(deck)
    sim2.ddcenergy(1)
    st_b = sim2.getState()
    assert np.array_equal(st["fx"], st_b["fx"]) and np.array_equal(st["fy"], st_b["fy"]) and np.array_equal(st["fz"], st_b["fz"])
    assert sim2.energyInfo().eion == e0.eion
    sim2.close()
    # energy bookkeeping across a list rebuild (updateRate = 20).  The generated membranes start from a lattice and relax
    # while they run, so this is a sanity bound (no blow-up, no lost pairs at the rebuild), not a drift bar: the drift bar
    # proper is tests/test_zz_nve_drift.py, against the reference's own 10k-step trace
    sim.nglf(nsteps)
    e1 = sim.energyInfo()
    assert e1.loop == nsteps and sim.lastListBuild() == 20 * (nsteps // 20)
    drift = abs((e1.eion + e1.rk) - (e0.eion + e0.rk))
    assert drift <= 5e-2 * e1.rk, (drift, e1.rk)
    # momentum: velocity-Verlet with internal forces keeps the total momentum (the decks start with none)
    sv = sim.getState()
    m = deck.array("specMass")[deck.array("species")]
    p = np.array([(m * sv[k]).sum() for k in ("vx", "vy", "vz")])
    assert np.abs(p).max() <= 1e-9 * (m * np.abs(sv["vx"])).sum()
    sim.close()
    return deck.n


@pytest.mark.skipif(not os.path.exists(REF_DUMP), reason="oracle/_ref/ref_dump not built")
@pytest.mark.parametrize("name", ["popc_100k", "ras_140k"])
def test_single_gpu_configs_against_live_reference(name):
    n = check_against_live_oracle(name)
    assert n > 80000


def test_bench_workload_properties():
    n = check_properties("membrane_1m")
    assert n > 1000000


@pytest.mark.skipif(not os.path.exists(REF_DUMP), reason="oracle/_ref/ref_dump not built")
def test_bench_workload_against_live_reference():
    """BASELINE.json's 1M-bead membrane, the deck bench.py times, against the unmodified reference run live on the same files
    (about half a minute of one host core): cells bit-exact, both pair lists as sets (through the order-independent hash),
    forces 1e-6, energies and virial 1e-9, and the energies after two integration steps."""
    n = check_against_live_oracle("membrane_1m", hashed=True)
    assert n > 1000000
