import sys, time; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np
import ddcmd_b200 as dd
from refdump import read_records
r = read_records('tests/golden/waterbox_ref.bin')
sim = dd.simulate_init('tests/golden/waterbox/object.data')
sim.ddcenergy(1)
e = sim.energyInfo()
u = r['units']
print('eion gpu', e.eion, 'ref', r['s0_energy'][0], 'rel', abs(e.eion-r['s0_energy'][0])/abs(r['s0_energy'][0]))
print('virial gpu', list(e.virial), 'ref', r['s0_energy'][6:12])
print('pairs listed', e.nPairsListed, 'ref', r['npairs'])
st = sim.getState()
f = np.stack([st['fx'],st['fy'],st['fz']]); fr = np.stack([r['s0_fx'],r['s0_fy'],r['s0_fz']])
print('force max abs err', np.abs(f-fr).max(), 'rel to rms', np.abs(f-fr).max()/np.sqrt((fr**2).mean()))
cell, dims, geom = sim.getCells()
print('dims', dims, r['geom_dims'], 'cells equal', np.array_equal(cell, r['cell']), 'geom', np.array_equal(geom, r['geom_parms'][:9]))
bi,bj,pr = sim.getPairs()
lab = r['s0_label']
ref = r['pairs0'].reshape(-1,2)
def key(a,b): return (np.minimum(a,b).astype(np.int64)<<32)|np.maximum(a,b).astype(np.int64)
print('pair sets equal', np.array_equal(np.sort(key(bi,bj)), np.sort(key(ref[:,0],ref[:,1]))), len(bi), len(ref))
print(sim.printinfo(e))
t=time.time(); sim.nglf(40); e2 = sim.energyInfo(); print('40 steps', time.time()-t)
tr = r['trace'].reshape(-1,16)
print('step40 eion', e2.eion, tr[39,1], 'rk', e2.rk, tr[39,2])
print(sim.printinfo(e2))
