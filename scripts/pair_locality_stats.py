"""Locality statistics of the Martini pair path, on the CPU (numpy + scipy): the evidence behind DESIGN.md section 3 "why the
pair kernel is a per-bead gather".

On a synthetic membrane in the engine's slot order (cell, 4x4x4 sub-cell Morton key) it measures
  1. the size of the union of the neighbor rows of T consecutive beads at j-cluster granularity C (what a cluster-pair /
     tile-list kernel would have to test per bead), and
  2. for the per-bead walk of the shipped kernel, the number of distinct 128-byte lines a warp touches per gather step when
     rows are in random order / sorted by slot, and the lane utilisation and line count of a "windowed" walk in which the
     lanes of a warp only consume entries of a common slot window.

    python scripts/pair_locality_stats.py            (about 2 minutes)
"""
import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
from ddcmd_b200 import synth
from scipy.spatial import cKDTree
s = synth.make_membrane(lx=300.0, ly=300.0, lz=130.0, seed=3)
L = np.array(s.box, float).reshape(-1)[[0,4,8]] if np.array(s.box).size==9 else np.array(s.box,float)
r = np.array(s.coords, float)
r = r - L*np.rint(r/L)        # in [-L/2, L/2]
n = len(r)
print("n", n, "box", L, "density", n/np.prod(L))
rl, rc = 15.0, 11.0
ncell = np.maximum(1, np.floor(L/rl)).astype(int)
d = L/ncell
q = (r + L/2)/d
ic = np.minimum(q.astype(int), ncell-1)
sub = np.minimum(((q-ic)*4).astype(int), 3)
def spread(v): return (v&1)|((v&2)<<2)
key = spread(sub[:,0])|(spread(sub[:,1])<<1)|(spread(sub[:,2])<<2)
cell = ic[:,0]+ncell[0]*(ic[:,1]+ncell[1]*ic[:,2])
order = np.lexsort((np.arange(n), key, cell))
r = r[order]
slot_of = np.empty(n,int); slot_of[order]=np.arange(n)
tree = cKDTree(r + L/2, boxsize=L)
t=time.time()
pl = tree.query_pairs(rl, output_type='ndarray')
pc = tree.query_pairs(rc, output_type='ndarray')
print("pairs list/bead", len(pl)/n, "cut/bead", len(pc)/n, time.time()-t)
# full-direction arrays
I = np.concatenate([pl[:,0], pl[:,1]]); J = np.concatenate([pl[:,1], pl[:,0]])
for T in (1,2,4,8,16,32):
    for C in (1,4,8):
        tile = I//T; cl = J//C
        k = np.unique(tile.astype(np.int64)*(n//C+1)+cl)
        ntile = (n+T-1)//T
        U = len(k)/ntile           # clusters per tile
        print("tile %2d cluster %d: clusters/tile %.1f  tests/bead %.1f  (warp steps per bead %.2f)" % (T, C, U, U*C, U*C*T/32/T))

# ---- 2. per-bead walk: lines per gather step ----
I = np.concatenate([pl[:,0], pl[:,1]]); J = np.concatenate([pl[:,1], pl[:,0]])
o = np.lexsort((J, I)); I=I[o]; J=J[o]
start = np.searchsorted(I, np.arange(n+1))
cnt = np.diff(start)
nw = n//32
rng = np.random.default_rng(0)
warps = rng.choice(nw, 300, replace=False)
# (a) lockstep over sorted rows: distinct 128B lines (4 slots) per step
lines_sorted=[]; lines_rand=[]; steps_lock=[]
for w in warps:
    rows=[J[start[i]:start[i+1]] for i in range(w*32,w*32+32)]
    m=max(len(x) for x in rows); steps_lock.append(m)
    for k in range(m):
        js=np.array([x[k] for x in rows if k<len(x)])
        lines_sorted.append(len(np.unique(js//4)))
    rr=[rng.permutation(x) for x in rows]
    for k in range(0,m,4):
        js=np.array([x[k] for x in rr if k<len(x)])
        lines_rand.append(len(np.unique(js//4)))
print("avg row", cnt.mean(), "lockstep steps/warp", np.mean(steps_lock))
print("lockstep sorted rows: lines/step", np.mean(lines_sorted), " random-order rows:", np.mean(lines_rand))
# (b) windowed walk
for W in (16,32,64,128,256):
    tot_steps=[]; tot_lines=[]; nwin=[]
    for w in warps:
        rows=[J[start[i]:start[i+1]] for i in range(w*32,w*32+32)]
        allw=np.unique(np.concatenate(rows)//W)
        st=0; ln=0
        for win in allw:
            per=[x[(x//W)==win] for x in rows]
            m=max(len(x) for x in per)
            st+=m
            for k in range(m):
                js=np.array([x[k] for x in per if k<len(x)])
                ln+=len(np.unique(js//4))
        tot_steps.append(st); tot_lines.append(ln); nwin.append(len(allw))
    print("W=%3d: windows/warp %.1f steps/warp %.1f (util %.2f) lines/warp %.0f lines/step %.1f" % (W, np.mean(nwin), np.mean(tot_steps), cnt.mean()/np.mean(tot_steps), np.mean(tot_lines), np.mean(tot_lines)/np.mean(tot_steps)))
