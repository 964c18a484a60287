"""Multi-GPU NGLFCONSTRAINT worker (run under torchrun): LANGEVIN groups, velocity constraints and the barostat on several ranks
against the single-rank reference traces of tests/golden/nglfc.npz - the per-bead random streams must follow migrating beads, the
barostat must see the global virial, constraint clusters are solved on their molecule's owner.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/mgpu_nglfc_worker.py <deck> <variant> [lx ly lz]
"""
import os
import sys
import tempfile

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ddcmd_b200 as dd  # noqa: E402
import nglfc_decks  # noqa: E402

if os.environ.get("DDCB200_TEST_EMU") == "1":
    import ctypes
    sys.path.insert(0, os.path.join(ROOT, "tests", "cpu_emu"))
    import build_emu
    dd._lib = dd._declare(ctypes.CDLL(os.environ.get("DDCB200_EMU_LIB") or build_emu.build()))


def main():
    deck, variant = sys.argv[1], sys.argv[2]
    lattice = tuple(int(x) for x in sys.argv[3:6]) if len(sys.argv) >= 6 else None
    nsteps = int(os.environ.get("DDCB200_NGLFC_STEPS", "40"))
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get("LOCAL_RANK", rank))
    ident = [dd.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    golden = os.path.join(ROOT, "tests", "golden")
    g = np.load(os.path.join(golden, "nglfc.npz"))
    key = "%s_%s_" % (deck, variant)
    tr, box = g[key + "trace"], g[key + "box"]
    lang, baro = nglfc_decks.VARIANTS[variant]
    tmp = [tempfile.mkdtemp(prefix="mgpu_nglfc_") if rank == 0 else None]
    dist.broadcast_object_list(tmp, src=0)
    d = [nglfc_decks.make_variant(golden, deck, variant, tmp[0]) if rank == 0 else None]
    dist.broadcast_object_list(d, src=0)
    sim = dd.simulate_init(os.path.join(d[0], "object.data"), device=local, rank=rank, nranks=world, lattice=lattice, nccl_id=ident[0])
    cons = int(sim.deck.s.nCons) > 0
    etol = 1e-9 if cons else 1e-10
    sim.ddcenergy(1)
    for s in range(nsteps):
        sim.eval_integrator(1)
        e = sim.energyInfo()
        assert e.loop == int(tr[s, 0])
        assert e.nPairsListed == int(tr[s, 14]), "pairs listed at loop %d" % e.loop
        etot = tr[s, 1] + tr[s, 2]
        assert abs((e.eion + e.rk) - etot) <= etol * max(abs(etot), abs(tr[s, 2])), ("Etot at loop %d" % e.loop, e.eion + e.rk, etot)
        assert abs(e.rk - tr[s, 2]) <= etol * abs(tr[s, 2]), "kinetic energy at loop %d" % e.loop
        h = sim.getBox()
        assert np.allclose([h[0], h[4], h[8]], box[s], rtol=1e-12, atol=0), "box at loop %d" % e.loop
    if lang and nsteps == len(tr):
        rng = sim.getRandom()
        assert np.array_equal(rng, g[key + "rng"]), "random streams"
    nfail = sim.constraintFailures()
    assert nfail == 0
    if rank == 0:
        import shutil
        shutil.rmtree(tmp[0], ignore_errors=True)
        print("MGPU_NGLFC_OK %s %s world=%d steps=%d" % (deck, variant, world, nsteps), flush=True)
    sim.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
