"""The wavelet compression protocol of the reference's post-processing unit test (BASELINE config 5):
LIB/POSTPROCESSING/post_compression_unit_test.f90:107-215.  For every threshold eps of the sweep:

  1) data on the equidistant grid of level Jmax: a Gauss blob, one component (set_block_testing_data, LIB/MESH/module_mesh.f90:89-155)
  2) sync_ghosts_tree + adapt_tree (coarsening indicator "threshold-state-vector", full wavelet transformation; coarse extension and
     security zone for lifted wavelets, neither for CDFX0)                                   -> Nb, the number of blocks kept
  3) sync_ghosts_tree + refineToEquidistant_tree (LIB/MESH/adaptToLevel_tree.f90:4-96) back to level Jmax
  4) relative L2 / Linfty error against the analytic field, summed / maximised over ranks

Everything heavy runs on the device(s) through the C ABI (wgpu_fwt / wgpu_threshold / wgpu_coarse_extension / wgpu_coarsen / wgpu_iwt_ce /
wgpu_refine behind WabbitGPU.adapt_tree, refine_tree and their multi-rank counterparts in DistributedWabbit)."""
from __future__ import annotations

import time
from typing import Optional, Sequence

import numpy as np

# the 51 thresholds of post_compression_unit_test.f90:107-120 (10^(-10 + k/5) printed with nine digits)
EPS_SWEEP = np.array([
    1.00000000e-10, 1.58489319e-10, 2.51188643e-10, 3.98107171e-10, 6.30957344e-10, 1.00000000e-09, 1.58489319e-09, 2.51188643e-09,
    3.98107171e-09, 6.30957344e-09, 1.00000000e-08, 1.58489319e-08, 2.51188643e-08, 3.98107171e-08, 6.30957344e-08, 1.00000000e-07,
    1.58489319e-07, 2.51188643e-07, 3.98107171e-07, 6.30957344e-07, 1.00000000e-06, 1.58489319e-06, 2.51188643e-06, 3.98107171e-06,
    6.30957344e-06, 1.00000000e-05, 1.58489319e-05, 2.51188643e-05, 3.98107171e-05, 6.30957344e-05, 1.00000000e-04, 1.58489319e-04,
    2.51188643e-04, 3.98107171e-04, 6.30957344e-04, 1.00000000e-03, 1.58489319e-03, 2.51188643e-03, 3.98107171e-03, 6.30957344e-03,
    1.00000000e-02, 1.58489319e-02, 2.51188643e-02, 3.98107171e-02, 6.30957344e-02, 1.00000000e-01, 1.58489319e-01, 2.51188643e-01,
    3.98107171e-01, 6.30957344e-01, 1.00000000e+00])

DOMAIN = 2.0            # params%domain_size (post_compression_unit_test.f90:83)
SIGMA0 = 0.3 / 15.0     # module_mesh.f90:104
AMPLI = 4.0


def compression_params(wavelet: str, Bs: int, Jmax: int, dim: int = 3, n_eqn: int = 1):
    """the fixed parameters of post_compression_unit_test.f90:83-104: domain 2, one component, Jmin = 1, coarse extension and security zone
    exactly for lifted wavelets, no dealiasing"""
    from .params import Params
    X, Y = int(wavelet[3]), int(wavelet[4])
    lifted = Y != 0
    p = Params(dim=dim, domain=(DOMAIN,) * 3, Bs=(Bs, Bs, Bs if dim == 3 else 1), wavelet=wavelet, g=X - 1 + max(Y - 1, 0), g_rhs=2, n_eqn=n_eqn,
               Jmax=Jmax, discretization="FD_4th_central")
    p.useCoarseExtension = 1 if lifted else 0
    p.useSecurityZone = 1 if lifted else 0
    return p.finalize()


def _wrap(x, xp):
    x = xp.where(x < -DOMAIN / 2.0, x + DOMAIN, x)
    return xp.where(x > DOMAIN / 2.0, x - DOMAIN, x)


def set_block_testing_data(Bs: int, level: np.ndarray, ixyz: np.ndarray, xp=np, device=None):
    """set_block_testing_data (module_mesh.f90:130-152) for the blocks listed: interiors [n, Bs, Bs, Bs] (z, y, x).  x = i*dx + x0 - L/2
    with x0 = ixyz*Bs*dx (get_block_spacing_origin), wrapped into [-L/2, L/2]; u = 1 + 4 exp(-(x^2 + y^2 + z^2) / (2 sigma0^2)).
    xp = numpy (host, the parity tests) or torch (device, the benchmark)."""
    if xp is np:
        lv = np.asarray(level, dtype=np.float64)
        dx = (2.0 ** (-lv)) * DOMAIN / float(Bs)
        idx = np.arange(Bs, dtype=np.float64)
        x0 = np.asarray(ixyz, dtype=np.float64) * float(Bs) * dx[:, None]
        ax = [_wrap(idx[None, :] * dx[:, None] + x0[:, a:a + 1] - DOMAIN / 2.0, np) for a in range(3)]
        X, Y, Z = ax[0][:, None, None, :], ax[1][:, None, :, None], ax[2][:, :, None, None]
        return 1.0 + AMPLI * np.exp(-((X ** 2 + Y ** 2) + Z ** 2) / (2.0 * SIGMA0 ** 2))
    torch = xp
    lv = torch.as_tensor(np.asarray(level, dtype=np.float64), device=device)
    dx = (2.0 ** (-lv)) * DOMAIN / float(Bs)
    idx = torch.arange(Bs, dtype=torch.float64, device=device)
    x0 = torch.as_tensor(np.asarray(ixyz, dtype=np.float64), device=device) * float(Bs) * dx[:, None]
    ax = [_wrap(idx[None, :] * dx[:, None] + x0[:, a:a + 1] - DOMAIN / 2.0, torch) for a in range(3)]
    X, Y, Z = ax[0][:, None, None, :], ax[1][:, None, :, None], ax[2][:, :, None, None]
    return 1.0 + AMPLI * torch.exp(-((X ** 2 + Y ** 2) + Z ** 2) / (2.0 * SIGMA0 ** 2))


def refineToEquidistant_tree(drv, forest, level: Optional[int] = None):
    """refineToEquidistant_tree (adaptToLevel_tree.f90:54-91), refinement branch: while a block is below `level`, flag every such block,
    refine (ghost synchronisation and gradedness are part of refine_tree here).  `drv`: WabbitGPU (one rank, takes and returns the forest)
    or DistributedWabbit (its own forest; flags in the global space-filling-curve order)."""
    distributed = hasattr(drv, "world")
    f = drv.forest if distributed else forest
    level = f.Jmax if level is None else level
    while True:
        lv = np.concatenate([f.active(r)[1] for r in range(f.n_ranks)])
        if lv.min() >= level:
            return f
        flags = (lv < level).astype(np.int32)
        f = drv.refine_tree(flags) if distributed else drv.refine_tree(f, flags)


class CompressionTest:
    """post_compression_unit_test on one GPU (sol = WabbitGPU) or across ranks (drv = DistributedWabbit).  The analytic field is evaluated on
    the device straight into / against the resident hvy_block (torch elementwise float64 on the library's device pointer)."""

    def __init__(self, sol, forest, drv=None):
        import torch
        self.sol, self.drv, self.torch = sol, drv, torch
        self.rank = drv.rank if drv is not None else 0
        self.forest0 = forest
        p = sol.params
        self.Bs, self.Jmax = p.Bs[0], forest.Jmax
        assert p.dim == 3 and p.n_eqn == 1
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.forest = forest
        self.stream = torch.cuda.ExternalStream(sol.stream) if sol.stream else torch.cuda.current_stream()

    def _forest(self):
        return self.drv.forest if self.drv is not None else self.forest

    def _hvy_block(self):
        """the resident hvy_block as a torch view; fetched anew every time: refinement and block moves write into the second buffer, which
        then BECOMES hvy_block (the device pointer changes)"""
        import ctypes as C
        from .multi import _DevPtr
        sol, Bs = self.sol, self.Bs
        ptr, n = C.c_void_p(), C.c_int64()
        sol._check(sol._lib.wgpu_device_pointer(sol._ctx, 0, 0, C.byref(ptr), C.byref(n)))
        return self.torch.as_tensor(_DevPtr(ptr.value, n.value), device=self.dev).view(sol.max_blocks, Bs, Bs, Bs)

    def _each_chunk(self, fn, chunk=4096):
        hvy, lvl, ixyz, _ = self._forest().active(self.rank)
        assert (np.diff(hvy) == 1).all() if len(hvy) > 1 else True
        U = self._hvy_block()
        with self.torch.cuda.stream(self.stream):
            for s0 in range(0, len(hvy), chunk):
                e = min(s0 + chunk, len(hvy))
                exact = set_block_testing_data(self.Bs, lvl[s0:e], ixyz[s0:e], self.torch, self.dev)
                fn(U[int(hvy[s0]) - 1:int(hvy[s0]) - 1 + (e - s0)], exact)

    def create_data(self):
        self._each_chunk(lambda blk, exact: blk.copy_(exact))
        self.sol._check(self.sol._lib.wgpu_synchronize(self.sol._ctx))

    def errors(self):
        acc = self.torch.zeros(4, dtype=self.torch.float64, device=self.dev)

        def fn(blk, exact):
            d = blk - exact
            acc[0] += (d * d).sum()
            acc[1] += (exact * exact).sum()
            acc[2] = self.torch.maximum(acc[2], d.abs().max())
            acc[3] = self.torch.maximum(acc[3], exact.abs().max())
        self._each_chunk(fn)
        a = acc.cpu().numpy()
        if self.drv is not None and self.drv.world > 1:
            s = self.drv.tr.allreduce_sum_np(a[:2].copy())
            m = self.drv.tr.allreduce_max_np(a[2:].copy())
            a = np.concatenate([s, m])
        return float(np.sqrt(a[0]) / np.sqrt(a[1])), float(a[2] / a[3])

    def run(self, eps_list: Sequence[float], sync=None):
        """Returns one record per eps: eps, Nb (blocks after adapt_tree), err_L2, err_Linfty, ms_adapt, ms_refine."""
        sol, drv = self.sol, self.drv
        sync = sync or (lambda: sol._check(sol._lib.wgpu_synchronize(sol._ctx)))
        self.forest = self.forest0
        out = []
        for eps in eps_list:
            f = self._forest()
            assert f.is_uniform and f.n_blocks == 8 ** self.Jmax
            self.create_data()
            sync()
            t0 = time.perf_counter()
            if drv is not None:
                _, n0, nb = drv.adapt_tree(eps=float(eps), Jmin=1, full_tree=True)
            else:
                self.forest, n0, nb = sol.adapt_tree(self.forest, eps=float(eps), Jmin=1, full_tree=True)
            sync()
            t1 = time.perf_counter()
            f = refineToEquidistant_tree(drv if drv is not None else sol, None if drv is not None else self.forest, self.Jmax)
            if drv is None:
                self.forest = f
            sync()
            t2 = time.perf_counter()
            e2, einf = self.errors()
            out.append({"eps": float(eps), "Nb": int(nb), "err_L2": e2, "err_Linfty": einf, "ms_adapt": round((t1 - t0) * 1e3, 2),
                        "ms_refine": round((t2 - t1) * 1e3, 2)})
        return out
