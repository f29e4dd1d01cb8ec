"""Two RK4 steps on 4096 blocks (Bs = 16, level 4) with the in-kernel mask of a translating sphere: the launches ncu captures for
profiles/r04_stage_kernel_sphere.txt"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wabbit_b200 import Forest, Params, WabbitGPU  # noqa: E402

p = Params(dim=3, domain=(6.283185307179586,) * 3, Bs=(16,) * 3, wavelet="CDF40", g=3, g_rhs=2, n_eqn=4, Jmax=4, discretization="FD_4th_central",
           skew_symmetry=True, c0=10.0, nu=3.125e-3, gamma_p=0.0, CFL=1.0, u_mean_set=(0.0, 0.0, 0.0), time_max=1.0e9, penalization=True,
           C_eta=1.0e-3).finalize()
forest = Forest.uniform(3, 4)
sol = WabbitGPU(p, max_blocks=forest.n_blocks)
sol.set_forest(forest)
host = np.zeros(sol.host_shape())
rng = np.random.default_rng(0)
host[:, :3] = 0.1 * rng.standard_normal(host[:, :3].shape)
sol.upload(host)
dx = p.domain[0] / (2 ** 4 * 16)
sol.set_mask_sphere((3.0, 3.1, 3.2), (0.5, 0.3, -0.2), 0.8, 1.5 * dx)
t, it = 0.0, 0
for _ in range(2):
    t, it, dt = sol.timeStep_tree(t, it)
sol.synchronize()
print("ok", t, dt)
