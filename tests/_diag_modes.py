import sys; sys.path.insert(0, '.')
import numpy as np
from bepuphysics2_b200 import scenes
from tests import util
names = ['qx','qy','qz','qw','px','py','pz','lx','ly','lz','ax','ay','az'] + ['L%d'%i for i in range(7)] + ['W%d'%i for i in range(7)]
for mode in (1, 2):
    for substeps in (1, 2):
        integ = util.bp.IntegratorDesc.default(); integ.angular_integration_mode = mode
        sc = scenes.shape_pile(200, seed=9)
        a = util.make_sim(sc, substeps=substeps, velocity_iterations=1, integrator=integ)
        b = util.make_sim(sc, substeps=substeps, velocity_iterations=1, integrator=integ)
        ra = util.run_oracle(a, 1/60); rb = util.run_gpu(b, 1/60, strict=True)
        x = ra['bodies'][:, util.MEANINGFUL]; y = rb['bodies'][:, util.MEANINGFUL]
        d = np.abs(x - y).max(axis=0)
        print('mode', mode, 'substeps', substeps, {n: float(v) for n, v in zip(names, d) if v > 0})
