"""Time the FWT kernel (CDF44, Bs = 16, 32768 blocks x 4 components) alone: ms per launch for the library currently in wabbit_b200/."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wabbit_b200 import Forest, Params, WabbitGPU  # noqa: E402
from wabbit_b200.solver import HVY_BLOCK, HVY_TMP  # noqa: E402

wavelet = sys.argv[1] if len(sys.argv) > 1 else "CDF44"
J = int(sys.argv[2]) if len(sys.argv) > 2 else 5
BS = int(sys.argv[3]) if len(sys.argv) > 3 else 16
p = Params(dim=3, domain=(6.283185307179586,) * 3, Bs=(BS,) * 3, wavelet=wavelet, g=6, g_rhs=2, n_eqn=4, Jmax=J, discretization="FD_4th_central",
           skew_symmetry=True, c0=10.0, nu=3.125e-3, gamma_p=0.0, CFL=1.0, u_mean_set=(0.0, 0.0, 0.0), time_max=1.0e9).finalize()
forest = Forest.uniform(3, J)
stream = torch.cuda.current_stream()
sol = WabbitGPU(p, max_blocks=forest.n_blocks, stream=stream.cuda_stream)
sol.set_forest(forest)
sol.setup_wavelet(wavelet)
host = np.random.default_rng(0).standard_normal((256,) + sol.host_shape()[1:])
first = np.arange(1, 257, dtype=np.int32)
sol.upload(host, hvy_ids=first)                     # host arrays are indexed by hvy id: fill 256 blocks, replicate them on the device
for s in range(256, forest.n_blocks, 256):
    sol.move_blocks(first, first + s)
for inverse in (False, True):
    f = sol.waveletReconstruction_tree if inverse else sol.waveletDecomposition_tree
    for _ in range(3):
        f((HVY_BLOCK, 0), (HVY_TMP, 0))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(20):
        f((HVY_BLOCK, 0), (HVY_TMP, 0))
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    B = 8 * 4 * ((BS + 12) ** 3 + BS ** 3) + 32
    print(f"{wavelet} Bs={BS} {'IWT' if inverse else 'FWT'} {forest.n_blocks} blocks: {ms:.3f} ms per launch, {B * forest.n_blocks / ms / 1e6:.0f} GB/s algorithmic", flush=True)
sol.close()
