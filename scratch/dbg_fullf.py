import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
import pic1dp_b200 as P
from helpers import OracleRun, copy_state, make_params, rel_err, synth_markers
op, gp = make_params(nx=256, capacity=60000, deltaf=0, iptcldist=0, density=[1.0], v0=[0.0])
st = synth_markers(op, 60000, seed=20)
ref = OracleRun(op, [[copy_state(st)]])
ref.init_field()
g = P.Pic1dGpu(gp)
g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
g.collect_charge(); g.solve_field()
f=g.get_field()
print('init rho', rel_err(f['chargeden'], ref.rho), 'E', rel_err(f['electric'], ref.E))
for it in range(2):
  for irk in (1,2):
    ref.push(irk); g.push(irk)
    out=g.get_markers(0)
    r=ref.st[0][0]
    xr=r['x'].copy(); ref.o.shape(xr)
    print(it,irk,'x eq',np.array_equal(out['x'],xr),'v eq',np.array_equal(out['v'],r['v']),'p eq',np.array_equal(out['p'],r['p']), 'maxdx', np.abs(out['x']-xr).max(), 'maxdv', np.abs(out['v']-r['v']).max())
    ref.collect_charge(); g.collect_charge()
    f=g.get_field(); print('   rho', rel_err(f['chargeden'], ref.rho), (f['chargeden']-ref.rho)[:4])
    ref.solve_field(); g.solve_field()
    f=g.get_field(); print('   E', rel_err(f['electric'], ref.E))
