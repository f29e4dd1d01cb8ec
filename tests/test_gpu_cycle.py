"""The adaptive cycle of the reference's time loop (LIB/MAIN/main.f90:314-443; the protocol of performance_test.f90):
adapt_tree (coarsening by wavelet thresholding) -> refine_tree("everywhere") -> RungeKuttaGeneric on the graded grid -> adapt_tree,
on the GPU against the same sequence assembled from the oracle's per-block routines (ghost synchronisation with level jumps,
waveletDecomposition_optimized_block, threshold_block, refineBlock, RHS_3D_acm, Runge-Kutta).  The light-data decisions
(completeness, gradedness) are the same host function on both sides, so block lists must be identical and heavy data bit-exact
after adapt / refine, 1e-12 after a time step."""
import numpy as np
import pytest

import oracle as O
from wabbit_b200 import Forest, WabbitGPU
from wabbit_b200.forest import coarsening_groups

from util import orc_grid, orc_params, relerr, tg_params

pytestmark = pytest.mark.gpu
Bs = 16


def keys(forest):
    hvy, lvl, ixyz, _ = forest.active(0)
    return [(int(l), int(a), int(b), int(c)) for l, (a, b, c) in zip(lvl, ixyz)], hvy


def to_array(po, grid_keys, data):
    u = np.zeros((len(grid_keys), 4, Bs + 2 * po.g, Bs + 2 * po.g, Bs + 2 * po.g))
    g = po.g
    for k, key in enumerate(grid_keys):
        u[k, :, g:g + Bs, g:g + Bs, g:g + Bs] = data[key]
    return u


def orc_adapt(w, po, forest, data, eps, Jmin):
    ks, _ = keys(forest)
    grid = orc_grid(forest)
    nbr = forest.neighbors(0)[:, :grid.n]
    u = to_array(po, ks, data)
    norm = O.norm_linfty_tree(po, u)
    norm[norm <= 1e-9] = 1.0
    O.sync_ghosts_leaf(grid, po, u, nbr, po.g, po.g, w.X, False)
    wd = np.zeros_like(u)
    O.fwt_tree(w, po, u, wd)
    st, _ = O.threshold_tree(po, wd, grid.level, eps, norm, "Linfty", None, level_ref=forest.Jmax)
    st = coarsening_groups(forest, st, Jmin)
    g, h = po.g, Bs // 2
    out = {}
    for k, key in enumerate(ks):
        if st[k] != -1:
            out[key] = data[key]
        else:
            L, x, y, z = key
            m = (L - 1, x // 2, y // 2, z // 2)
            blk = out.setdefault(m, np.zeros((4, Bs, Bs, Bs)))
            qx, qy, qz = x % 2, y % 2, z % 2
            blk[:, qz * h:(qz + 1) * h, qy * h:(qy + 1) * h, qx * h:(qx + 1) * h] = wd[k][:, g:g + Bs:2, g:g + Bs:2, g:g + Bs:2]
    return out, st


def orc_refine(w, po, forest, data):
    ks, _ = keys(forest)
    grid = orc_grid(forest)
    nbr = forest.neighbors(0)[:, :grid.n]
    u = to_array(po, ks, data)
    O.sync_ghosts_leaf(grid, po, u, nbr, po.g, po.g, w.X, False)
    g = po.g
    out = {}
    for k, (L, x, y, z) in enumerate(ks):
        if L >= forest.Jmax:
            out[(L, x, y, z)] = data[(L, x, y, z)]
            continue
        d = O.refine_block(w.X, po, u[k])
        for q in range(8):
            qq = ((q >> 1) & 1, q & 1, (q >> 2) & 1)
            out[(L + 1, 2 * x + qq[0], 2 * y + qq[1], 2 * z + qq[2])] = d[q][:, g:g + Bs, g:g + Bs, g:g + Bs].copy()
    return out


def gpu_data(sol, forest, po):
    ks, hvy = keys(forest)
    got = np.zeros(sol.host_shape())
    sol.download(got, g_sync=0)
    g = po.g
    return {key: got[h - 1][:, g:g + Bs, g:g + Bs, g:g + Bs].copy() for key, h in zip(ks, hvy)}


def blob_field(po, forest):
    """Taylor-Green plus a Gaussian vortex blob: smooth almost everywhere, sharp in one corner of the domain"""
    ks, _ = keys(forest)
    data = {}
    L0 = 6.283185307179586
    for (L, bx, by, bz) in ks:
        dx = L0 / (2 ** L * Bs)
        x = (bx * Bs + np.arange(Bs)) * dx
        y = (by * Bs + np.arange(Bs)) * dx
        z = (bz * Bs + np.arange(Bs)) * dx
        Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
        r2 = (X - 1.2) ** 2 + (Y - 1.9) ** 2 + (Z - 1.4) ** 2
        blob = np.exp(-r2 / (2 * 0.15 ** 2))
        data[(L, bx, by, bz)] = np.stack([np.sin(X) * np.cos(Y) * np.cos(Z) + 2.0 * blob, -np.cos(X) * np.sin(Y) * np.cos(Z) - blob,
                                          0.5 * blob, (np.cos(2 * X) + np.cos(2 * Y)) * (np.cos(2 * Z) + 2.0) / 16.0])
    return data


def test_adapt_refine_step_adapt_cycle():
    w = O.setup_wavelet("CDF40")
    Jmax, eps, Jmin = 4, 1.0e-2, 1
    p = tg_params(Bs=Bs, J=Jmax, wavelet_g=w.g_default)
    p.wavelet = "CDF40"
    po = orc_params(p)
    forest = Forest.uniform(3, 3, Jmax=Jmax)
    sol = WabbitGPU(p, max_blocks=4096)
    sol.setup_wavelet("CDF40")
    sol.set_forest(forest)
    data = blob_field(po, forest)
    ks, hvy = keys(forest)
    host = np.zeros(sol.host_shape())
    g = po.g
    for key, h in zip(ks, hvy):
        host[h - 1][:, g:g + Bs, g:g + Bs, g:g + Bs] = data[key]
    sol.upload(host)

    def same(a, b, exact=True):
        assert sorted(a) == sorted(b)
        for k in a:
            if exact:
                assert np.array_equal(a[k], b[k]), k
            else:
                assert relerr(a[k], b[k]) <= 1e-12, k

    # two coarsening sweeps
    sizes = [forest.n_blocks]
    for _ in range(2):
        ref_forest = forest
        forest, n0, n1 = sol.adapt_tree(forest, eps=eps, Jmin=Jmin)
        data, _ = orc_adapt(w, po, ref_forest, data, eps, Jmin)
        same(gpu_data(sol, forest, po), data)
        sizes.append(n1)
    assert sizes[1] < sizes[0] and not forest.is_uniform
    # refine everywhere, one RK4 step on the graded grid, coarsen again
    ref_forest = forest
    forest = sol.refine_tree(forest)
    data = orc_refine(w, po, ref_forest, data)
    same(gpu_data(sol, forest, po), data)
    dt = sol.RungeKuttaGeneric(0.0, 0)
    ks, _ = keys(forest)
    grid = orc_grid(forest)
    nbr = forest.neighbors(0)[:, :grid.n]
    u = to_array(po, ks, data)
    work = [np.zeros_like(u) for _ in range(5)]
    dt_ref = O.rk_generic(grid, po, u, work, 0.0, sync=lambda h: O.sync_ghosts_leaf(grid, po, h, nbr, po.g_rhs, po.g_rhs, w.X, False))
    assert dt == dt_ref
    data = {key: u[k][:, g:g + Bs, g:g + Bs, g:g + Bs].copy() for k, key in enumerate(ks)}
    same(gpu_data(sol, forest, po), data, exact=False)
    n_before = forest.n_blocks
    forest, n0, n1 = sol.adapt_tree(forest, eps=eps, Jmin=Jmin)
    assert n1 < n_before
    print("adaptive cycle block counts:", sizes, "-> refined", n_before, "-> adapted", n1)
    sol.close()
