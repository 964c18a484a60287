"""Multi-GPU parity worker (run under torchrun, one rank per GPU): the ddc-decomposed CUDA path against the
outputs of the unmodified single-rank reference CPU path (tests/golden/<deck>/ref.npz).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/mgpu_worker.py <deck> [lx ly lz]
"""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ddcmd_b200 as dd  # noqa: E402

if os.environ.get("DDCB200_TEST_EMU") == "1":
    # no-GPU container: run the same worker on the CPU emulation of the CUDA sources (tests/cpu_emu, test infrastructure)
    import ctypes
    sys.path.insert(0, os.path.join(ROOT, "tests", "cpu_emu"))
    import build_emu
    dd._lib = dd._declare(ctypes.CDLL(os.environ.get("DDCB200_EMU_LIB") or build_emu.build()))     # DDCB200_EMU_LIB: e.g. an ASan build


def gather_by_bead(sim, n, keys):
    st = sim.getState()
    beads = sim.getLocalBeads()
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, (beads, {k: st[k] for k in keys}))
    out = {k: np.full(n, np.nan) for k in keys}
    count = np.zeros(n, np.int32)
    for b, d in parts:
        count[b] += 1
        for k in keys:
            out[k][b] = d[k]
    assert np.all(count == 1), "every bead must be local on exactly one rank"
    return out, [len(b) for b, _ in parts]


def pairkey(a, b):
    a = a.astype(np.int64)
    b = b.astype(np.int64)
    return (np.minimum(a, b) << 32) | np.maximum(a, b)


def main():
    name = sys.argv[1]
    lattice = tuple(int(x) for x in sys.argv[2:5]) if len(sys.argv) >= 5 else None
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get("LOCAL_RANK", rank))
    ident = [dd.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    g = os.path.join(ROOT, "tests", "golden", name)
    ref = np.load(os.path.join(g, "ref.npz"))
    sim = dd.simulate_init(os.path.join(g, "object.data"), device=local, rank=rank, nranks=world, lattice=lattice, nccl_id=ident[0])
    n = sim.deck.n
    keys = ("rx", "ry", "rz", "vx", "vy", "vz", "fx", "fy", "fz")

    # ---- step 0: forces, energies, virial, pair membership ----
    sim.ddcenergy(1)
    e = sim.energyInfo()
    st, nloc = gather_by_bead(sim, n, keys)
    bi, bj, pr = sim.getPairs()
    parts = [None] * world
    dist.all_gather_object(parts, (bi, bj, pr))
    if rank == 0:
        f = np.stack([st["fx"], st["fy"], st["fz"]], 1)
        fr = np.stack([ref["s0_fx"], ref["s0_fy"], ref["s0_fz"]], 1)
        rms = np.sqrt((fr ** 2).sum(1).mean())
        ferr = (np.sqrt(((f - fr) ** 2).sum(1)) / np.maximum(np.sqrt((fr ** 2).sum(1)), rms)).max()
        en = ref["s0_energy"]
        escale = max(abs(en[0]), 1e-3 * np.abs(ref["s0_fx"]).sum())
        assert ferr < 1e-6, ferr
        assert abs(e.eion - en[0]) <= 1e-9 * escale, (e.eion, en[0])
        assert np.allclose(np.array(e.virial[:]), en[6:12], rtol=1e-9, atol=1e-9 * np.abs(en[6:9]).max())
        assert e.nPairsListed == int(ref["npairs"][0]), (e.nPairsListed, int(ref["npairs"][0]))
        # the reference's ownership rule: a pair is reported by the rank where its smaller-gid bead is local
        abi = np.concatenate([p[0] for p in parts]); abj = np.concatenate([p[1] for p in parts]); apr = np.concatenate([p[2] for p in parts])
        p0 = ref["pairs0"].reshape(-1, 2); p1 = ref["pairs1"].reshape(-1, 2)
        assert np.array_equal(np.sort(pairkey(abi[apr == 0], abj[apr == 0])), np.sort(pairkey(p0[:, 0], p0[:, 1])))
        assert np.array_equal(np.sort(pairkey(abi[apr == 1], abj[apr == 1])), np.sort(pairkey(p1[:, 0], p1[:, 1])))
        # molecular virial / pressure: every molecule is summed once, on its owner (ghost copies elsewhere contribute nothing);
        # checked against a single-rank context of the same deck on this rank's device
        one = dd.simulate_init(os.path.join(g, "object.data"), device=local)
        one.ddcenergy(1)
        e1 = one.energyInfo()
        one.close()
        mv, mv1 = np.array(e.molVirial[:]), np.array(e1.molVirial[:])
        assert np.allclose(mv, mv1, rtol=1e-9, atol=1e-9 * np.abs(mv1).max()), (mv, mv1)
        assert abs(e.pMolecular - e1.pMolecular) <= 1e-9 * max(abs(e1.pMolecular), abs(e1.pion)), (e.pMolecular, e1.pMolecular)
        assert e.nMolecules == e1.nMolecules
        print("step0 ok: locals per rank %s, force err %.2e, eion %.12g, pMolecular %.12g" % (nloc, ferr, e.eion, e.pMolecular), flush=True)

    # ---- 40 steps: halo every step, re-domain + migration at steps 20 and 40 ----
    tr = ref["trace"].reshape(-1, 16)
    sim.nglf(20)
    e = sim.energyInfo()
    st, nloc = gather_by_bead(sim, n, keys)
    if rank == 0:
        etot = tr[19, 1] + tr[19, 2]
        assert abs((e.eion + e.rk) - etot) <= 1e-9 * max(abs(etot), abs(tr[19, 2])), (e.eion + e.rk, etot)
        assert abs(e.rk - tr[19, 2]) <= 1e-9 * abs(tr[19, 2])
        assert e.nPairsListed == int(tr[19, 14]), (e.nPairsListed, int(tr[19, 14]))
        assert np.abs(st["rx"] - ref["s20_rx"]).max() < 1e-9
        assert np.abs(st["vx"] - ref["s20_vx"]).max() < 1e-12
    sim.nglf(20)
    e = sim.energyInfo()
    st, nloc = gather_by_bead(sim, n, keys)
    if rank == 0:
        etot = tr[39, 1] + tr[39, 2]
        assert abs((e.eion + e.rk) - etot) <= 1e-9 * max(abs(etot), abs(tr[39, 2])), (e.eion + e.rk, etot)
        dz = np.abs(st["rz"] - ref["sN_rz"])
        assert np.quantile(dz, 0.99) < 1e-9 and dz.max() < 1e-6, (np.quantile(dz, 0.99), dz.max())
    # ---- a ddcMD-format restart written from the decomposed state: every bead once, in the deck's order ----
    import tempfile
    tmp = tempfile.mkdtemp(prefix="mgpu_snap_") if rank == 0 else None
    snap = sim.writeRestart(dirname=os.path.join(tmp, "snapshot.mgpu") if rank == 0 else None, restart_link=False)
    if rank == 0:
        import shutil
        raw = open(os.path.join(snap, "atoms#000000"), "rb").read()
        recs = raw[raw.index(b"\n\n", raw.index(b"}")) + 2:].split(b"\n")[:n]
        got = np.array([[float(x) for x in r.split()[5:8]] for r in recs])
        h = np.array([sim.getBox()[k] for k in (0, 4, 8)])
        d = got - np.stack([st["rx"], st["ry"], st["rz"]], 1) * dd.units_convert(1.0, None, "Angstrom")
        d -= h * dd.units_convert(1.0, None, "Angstrom") * np.rint(d / (h * dd.units_convert(1.0, None, "Angstrom")))
        assert np.abs(d).max() < 1e-10, np.abs(d).max()
        assert b"loop=40;" in raw[:2000]
        shutil.rmtree(tmp)
        print("MGPU_OK %s world=%d locals=%s" % (name, world, nloc), flush=True)
    sim.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
