"""End-to-end leg of bench.py (upload -> RK4 step -> download per step, page-locked host arrays) for every transfer mode
(wgpu_set_transfer_mode: copy engines vs zero-copy kernels, per direction) and 1..3 trees in flight.  Run on the GPU box:
    python profiles/r5_e2e_modes.py > gpurun_out/e2e_modes.log"""
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
from wabbit_b200 import Forest, WabbitGPU
from wabbit_b200 import solver as S

a = types.SimpleNamespace(level=int(os.environ.get("LEVEL", "5")), bs=16)
p = bench.make_params(a)
forest = Forest.uniform(3, a.level, block_dist="sfc_hilbert", n_ranks=1)
hvy, lvl, ixyz, _ = forest.active(0)
torch.cuda.set_device(0)
stream = torch.cuda.current_stream()
steps = 3
MODES = [(1, 1), (0, 0)] if os.environ.get('QUICK') else [(1, 1), (1, 0), (0, 1), (0, 0)]
for up, down in MODES:
    if True:
        orig = S.WabbitGPU.__init__

        def init(self, *args, _o=orig, **kw):
            _o(self, *args, **kw)
            self.set_transfer_mode(bool(up), bool(down))
        S.WabbitGPU.__init__ = init
        sol = WabbitGPU(p, max_blocks=forest.max_blocks, device=0, stream=stream.cuda_stream)
        sol.set_forest(forest, 0)
        shape = sol.host_shape()
        host = torch.empty(shape, dtype=torch.float64, pin_memory=True)
        bench.taylor_green_host(p, ixyz, lvl, host.numpy())
        t, it = 0.0, 0
        # phases of one sequential step
        for rep in range(2):
            torch.cuda.synchronize(); w0 = time.perf_counter()
            sol.upload_ptr(host.data_ptr(), shape[1], hvy_ids=hvy)
            torch.cuda.synchronize(); w1 = time.perf_counter()
            t, it, _ = sol.timeStep_tree(t, it)
            torch.cuda.synchronize(); w2 = time.perf_counter()
            sol.download_ptr(host.data_ptr(), shape[1], hvy_ids=hvy, g_sync=0)
            torch.cuda.synchronize(); w3 = time.perf_counter()
        gb = forest.n_blocks * 4 * a.bs ** 3 * 8 / 1e9
        print(f"up={'dma' if up else 'sm '} down={'dma' if down else 'sm '}  upload {1e3*(w1-w0):7.1f} ms ({gb/(w1-w0):5.1f} GB/s of interiors)  "
              f"step {1e3*(w2-w1):6.1f} ms  download {1e3*(w3-w2):7.1f} ms ({gb/(w3-w2):5.1f} GB/s)  sequential {forest.n_blocks/(w3-w0):9.0f} block-updates/s", flush=True)
        for trees in (2, 3, 4) if os.environ.get('QUICK') else (2, 3):
            v = bench.e2e_pipelined(a, p, forest, 0, sol, host, shape, hvy, steps, trees)
            print(f"      {trees} trees in flight: {v:9.0f} block-updates/s", flush=True)
        assert np.isfinite(host.numpy()[:4]).all()
        sol.close()
        S.WabbitGPU.__init__ = orig
        del host
