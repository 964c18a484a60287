"""CPU tests of the multi-GPU host logic (no GPU): the domain classification (ddcb200_ddcPlan = the same
__host__ __device__ predicates the re-domain kernels run: owner brick of the ownership bead, bounding boxes,
ghost test), checked single-process for its invariants - up to the 16 ranks the library accepts - and across
two real processes over gloo: what one rank lists as "mine, ghost on p" is what p lists as "ghost here, owned
by that rank" (DESIGN.md "Multi-GPU"; reference ddcSendRecvTables, src/ddcSendRecv.c:41-277).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import ddcmd_b200 as dd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _deck(golden_dir, name="ras_small"):
    deck = dd.Deck(os.path.join(golden_dir, name, "object.data"))
    rx, ry, rz = (deck.array(k).copy() for k in ("rx", "ry", "rz"))
    ob = np.arange(deck.n, dtype=np.int32)
    off, mb = deck.array("molOffset"), deck.array("molBeads")
    for m in range(len(off) - 1):
        ob[mb[off[m]:off[m + 1]]] = mb[off[m]]
    h = np.array(deck.s.params.h[:])
    rlist = deck.s.params.rmax + deck.s.params.deltaR
    return deck, h, rlist, rx, ry, rz, ob


def _min_image(d, L):
    return d - L * np.round(d / L)


@pytest.mark.parametrize("lattice", [(2, 1, 1), (2, 2, 1), (1, 1, 2), (2, 2, 2), (4, 4, 1)])
def test_plan_invariants(golden_dir, lattice):
    deck, h, rlist, rx, ry, rz, ob = _deck(golden_dir)
    n = deck.n
    nranks = lattice[0] * lattice[1] * lattice[2]
    plans = [dd.ddc_plan(h, lattice, rlist, rx, ry, rz, r, ob) for r in range(nranks)]
    owner = plans[0][0]
    for o, _ in plans[1:]:
        assert np.array_equal(o, owner)                         # every rank derives the same owners
    assert owner.min() >= 0 and owner.max() < nranks
    # molecules are whole on one rank (ddcRuleMolecule)
    mol = (deck.array("gid") >> np.uint64(32)).astype(np.int64)
    for m in np.unique(mol[ob != np.arange(n)]):
        assert len(np.unique(owner[mol == m])) == 1
    # each bead is local on exactly one rank
    local = np.stack([(m >> 31) & 1 for _, m in plans])
    assert np.all(local.sum(0) == 1)
    for r, (_, m) in enumerate(plans):
        assert np.array_equal(((m >> 31) & 1).astype(bool), owner == r)
    # send list r->p == recv list p<-r, as sets and in (ascending bead) order
    for r in range(nranks):
        for p in range(nranks):
            if p == r:
                continue
            send = np.nonzero(((plans[r][1] >> 31) & 1) & ((plans[r][1] >> p) & 1))[0]
            recv = np.nonzero(((plans[p][1] >> 30) & 1) & ((plans[p][1] >> r) & 1))[0]
            assert np.array_equal(send, recv)
    # completeness: every pair within the list range has, on the owner of either bead, the partner present
    L = np.array([h[0], h[4], h[8]])
    pos = np.stack([rx, ry, rz], 1)
    rng = np.random.default_rng(0)
    sample = rng.choice(n, size=min(n, 400), replace=False)
    for i in sample:
        d = _min_image(pos - pos[i], L)
        nb = np.nonzero((d ** 2).sum(1) < rlist ** 2)[0]
        m = plans[owner[i]][1]
        present = ((m >> 31) & 1).astype(bool) | (((m >> 30) & 1) != 0)
        assert np.all(present[nb]), "a neighbour within rcut+skin is neither local nor ghost on the owner"


def test_plan_ghost_fraction_is_a_shell(golden_dir):
    """Ghosts are a shell, not the world: with 2 bricks along the long axis of the RAS patch a rank
    must not import every foreign bead."""
    deck, h, rlist, rx, ry, rz, ob = _deck(golden_dir)
    owner, m = dd.ddc_plan(h, (1, 1, 2), rlist, rx, ry, rz, 0, ob)
    ghosts = (((m >> 30) & 1) != 0).sum()
    foreign = (owner != 0).sum()
    assert 0 < ghosts < foreign


WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import ddcmd_b200 as dd
from test_ddc_cpu import _deck
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
deck, h, rlist, rx, ry, rz, ob = _deck({golden!r})
lattice = dd.default_lattice(world, [h[0], h[4], h[8]])
owner, mask = dd.ddc_plan(h, lattice, rlist, rx, ry, rz, rank, ob)
# halo exchange over gloo using only locally derived lists: positions of my beads that the peer holds as ghosts
peer = 1 - rank
send = np.nonzero(((mask >> 31) & 1) & ((mask >> peer) & 1))[0]
recv = np.nonzero(((mask >> 30) & 1) & ((mask >> peer) & 1))[0]
# counts are NOT exchanged: the receive buffer is sized from the local plan alone
out = torch.from_numpy(np.stack([rx[send], ry[send], rz[send]], 1).copy())
inp = torch.empty((len(recv), 3), dtype=torch.float64)
reqs = [dist.isend(out, peer), dist.irecv(inp, peer)]
for r in reqs:
    r.wait()
got = inp.numpy()
assert np.array_equal(got, np.stack([rx[recv], ry[recv], rz[recv]], 1)), "ghost payload does not line up with the receiver's own list"
# every bead is owned exactly once across ranks
nloc = torch.tensor([int(((mask >> 31) & 1).sum())])
dist.all_reduce(nloc)
assert int(nloc) == deck.n
sys.stdout.write("RANKOK%d lattice=%s locals=%d send=%d recv=%d\n" % (rank, lattice, int(((mask >> 31) & 1).sum()), len(send), len(recv)))
sys.stdout.flush()
dist.destroy_process_group()
'''


def test_two_ranks_gloo_halo_lists_agree(golden_dir, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, golden=golden_dir))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29511", str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "RANKOK0" in r.stdout and "RANKOK1" in r.stdout, r.stdout
