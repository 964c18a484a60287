"""GPU parity tests: the CUDA path, called through the C-ABI, against the outputs of the
UNMODIFIED reference CPU path (tests/golden/*/ref.npz, produced by oracle/_ref/ref_dump).

Bars (BASELINE.json north_star): cell assignment and pair-list membership bit-exact;
per-bead forces within 1e-6 relative; total energy per step within 1e-9 relative.
"""
import os

import numpy as np
import pytest

import ddcmd_b200 as dd

pytestmark = pytest.mark.gpu

DECKS = ["waterbox", "popc_small", "ras_small", "tiny2"]   # tiny2: two cells per axis in x and y
F_TOL = 1e-6      # per-bead force, relative to max(|f_ref|, rms force)
E_TOL = 1e-9      # energies, relative


def _load(golden_dir, name):
    ref = np.load(os.path.join(golden_dir, name, "ref.npz"))
    sim = dd.simulate_init(os.path.join(golden_dir, name, "object.data"))
    return sim, ref


def _pairkey(a, b):
    a = a.astype(np.int64)
    b = b.astype(np.int64)
    return (np.minimum(a, b) << 32) | np.maximum(a, b)


def _force_err(st, ref, prefix):
    f = np.stack([st["fx"], st["fy"], st["fz"]], 1)
    fr = np.stack([ref[prefix + "fx"], ref[prefix + "fy"], ref[prefix + "fz"]], 1)
    rms = np.sqrt((fr ** 2).sum(1).mean())
    scale = np.maximum(np.sqrt((fr ** 2).sum(1)), rms)
    return (np.sqrt(((f - fr) ** 2).sum(1)) / scale).max()


@pytest.mark.parametrize("name", DECKS)
def test_cells_bit_exact(golden_dir, name):
    sim, ref = _load(golden_dir, name)
    sim.constructList()
    cell, dims, geom = sim.getCells()
    assert list(dims) == list(ref["geom_dims"][:3])
    assert np.array_equal(geom, ref["geom_parms"][:9])          # min, max, d: bitwise
    assert np.array_equal(cell, ref["cell"])


@pytest.mark.parametrize("name", DECKS)
def test_pair_membership_bit_exact(golden_dir, name):
    sim, ref = _load(golden_dir, name)
    sim.constructList()
    bi, bj, pr = sim.getPairs()
    p0 = ref["pairs0"].reshape(-1, 2)
    p1 = ref["pairs1"].reshape(-1, 2)
    assert len(bi) == int(ref["npairs"][0])
    got0 = np.sort(_pairkey(bi[pr == 0], bj[pr == 0]))
    got1 = np.sort(_pairkey(bi[pr == 1], bj[pr == 1]))
    assert np.array_equal(got0, np.sort(_pairkey(p0[:, 0], p0[:, 1])))     # interacting list (ifirst[0])
    assert np.array_equal(got1, np.sort(_pairkey(p1[:, 0], p1[:, 1])))     # pruned list (ifirst[1])
    # the owner of every pair is the smaller gid, as in pairlist1
    lab = ref["s0_label"]
    assert np.all(lab[bi] < lab[bj])


@pytest.mark.parametrize("name", DECKS)
def test_step0_forces_energy_virial(golden_dir, name):
    sim, ref = _load(golden_dir, name)
    sim.ddcenergy(1)
    e = sim.energyInfo()
    st = sim.getState()
    assert _force_err(st, ref, "s0_") < F_TOL
    en = ref["s0_energy"]
    escale = max(abs(en[0]), 1e-3 * np.abs(ref["s0_fx"]).sum())    # guards decks whose eion nearly cancels
    assert abs(e.eion - en[0]) <= E_TOL * escale
    vir = np.array(e.virial[:])
    assert np.allclose(vir, en[6:12], rtol=1e-9, atol=1e-9 * np.abs(en[6:9]).max())
    be = ref["bioEnergies"]   # bond, angle, ub, torsion, impr, ...
    assert np.isclose(e.eBond, be[0], rtol=1e-9, atol=1e-14)
    assert np.isclose(e.eAngle, be[1], rtol=1e-9, atol=1e-14)
    assert np.isclose(e.eTorsion, be[3], rtol=1e-9, atol=1e-14)
    assert np.isclose(e.eImproper, be[4], rtol=1e-9, atol=1e-14)


@pytest.mark.parametrize("name", DECKS)
def test_trajectory_40_steps(golden_dir, name):
    """nglf x 40 (two list rebuilds): energies at every 20th step and the final state."""
    sim, ref = _load(golden_dir, name)
    tr = ref["trace"].reshape(-1, 16)
    sim.nglf(20)
    e = sim.energyInfo()
    etot_ref = tr[19, 1] + tr[19, 2]
    assert abs((e.eion + e.rk) - etot_ref) <= 1e-9 * max(abs(etot_ref), abs(tr[19, 2]))
    assert abs(e.rk - tr[19, 2]) <= 1e-9 * abs(tr[19, 2])
    assert e.nPairsListed == int(tr[19, 14])
    st = sim.getState()
    assert np.abs(st["rx"] - ref["s20_rx"]).max() < 1e-9
    assert np.abs(st["vx"] - ref["s20_vx"]).max() < 1e-12
    sim.nglf(20)
    e = sim.energyInfo()
    etot_ref = tr[39, 1] + tr[39, 2]
    assert abs((e.eion + e.rk) - etot_ref) <= 1e-9 * max(abs(etot_ref), abs(tr[39, 2]))
    st = sim.getState()
    assert _force_err(st, ref, "sN_") < 1e-5     # chaotic growth of 1e-16 rounding differences over 40 steps
    dz = np.abs(st["rz"] - ref["sN_rz"])
    # a near-planar dihedral (acos near +-1) amplifies 1e-16 rounding differences to ~1e-8 on a few beads
    assert np.quantile(dz, 0.99) < 1e-9 and dz.max() < 1e-6


@pytest.mark.parametrize("name", ["waterbox", "popc_small", "ras_small"])
def test_deterministic_forces(golden_dir, name):
    """Neither the pair nor the bonded kernel uses a floating-point atomic: two runs (25 steps, one rebuild) are bitwise identical
    in forces, positions and every energy term - also on the decks with bonds, angles, dihedrals and restraints."""
    out = []
    for _ in range(2):
        sim, _ = _load(golden_dir, name)
        sim.ddcenergy(1)
        a = sim.getState()
        e0 = sim.energyInfo()
        sim.nglf(25)
        b = sim.getState()
        e1 = sim.energyInfo()
        out.append((a, b, [e0.eion, e0.eBond, e0.eAngle, e0.eTorsion, e0.eImproper, e0.eRestraint] + list(e0.virial[:]),
                    [e1.eion, e1.rk, e1.eBond, e1.eAngle] + list(e1.virial[:])))
        sim.close()
    for k in ("fx", "fy", "fz"):
        assert np.array_equal(out[0][0][k], out[1][0][k]), k
    for k in ("rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz"):
        assert np.array_equal(out[0][1][k], out[1][1][k]), k
    assert out[0][2] == out[1][2] and out[0][3] == out[1][3]


def test_printinfo_line_matches_reference_data_file(golden_dir):
    """Step-0 'data' line of examples/waterbox pinned by SURVEY.md (the reference's own output)."""
    sim, _ = _load(golden_dir, "waterbox")
    sim.ddcenergy(1)
    line = sim.printinfo().split()
    assert line[0] == "000000000000"
    assert abs(float(line[2]) - (-26.988954808928)) < 2e-12      # Etotal kJ/mol/bead
    assert abs(float(line[4]) - (-26.988954808928)) < 2e-12      # Epot
    assert abs(float(line[6]) - (-369.496497981831)) < 1e-8      # molecular pressure, bar
    assert abs(float(line[7]) - 133.942256177663) < 1e-9         # volume per bead


def _mgpu(nproc, port, name, lattice=()):
    import subprocess
    import sys
    if dd.lib().ddcb200_deviceCount() < nproc:
        pytest.skip("needs %d GPUs" % nproc)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % nproc, "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(root, "tests", "mgpu_worker.py"), name] + [str(x) for x in lattice]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MGPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-6000:]


@pytest.mark.parametrize("name", ["popc_small", "ras_small", "waterbox", "tiny2"])
def test_two_gpus_match_reference(name):
    """ddc decomposition over 2 GPUs (migration + ghost exchange every 20 steps, NCCL halo per step) against the single-rank reference."""
    _mgpu(2, 29533, name)


@pytest.mark.parametrize("name,lattice", [("popc_small", (2, 2, 1)), ("ras_small", (4, 1, 1))])
def test_four_gpus_match_reference(name, lattice):
    _mgpu(4, 29534, name, lattice)


@pytest.mark.parametrize("name,lattice", [("popc_small", (4, 2, 1)), ("ras_small", (2, 2, 2))])
def test_eight_gpus_match_reference(name, lattice):
    """4x2x1: the lattice of the 1M-bead scaling run (some rank pairs exchange nothing); 2x2x2: bricks in every direction."""
    _mgpu(8, 29535, name, lattice)
