"""BASELINE config 5 at test size: the compression protocol of post_compression_unit_test.f90:149-215 -- equidistant grid at Jmax,
Gaussian blob 1 + 4 exp(-r^2 / (2 sigma^2)) (set_block_testing_data, LIB/MESH/module_mesh.f90:89-130), adapt_tree until the grid is
stationary, number of blocks, refinement back to the equidistant grid, error against the analytic field -- on the GPU against the
same sequence assembled from the oracle's per-block routines, for a sweep of thresholds (CDF40, eps_norm = Linfty, normalised)."""
import numpy as np
import pytest

import oracle as O
from wabbit_b200 import Forest, WabbitGPU

from test_gpu_cycle import Bs, gpu_data, keys, orc_adapt, orc_refine
from util import orc_params, tg_params

pytestmark = pytest.mark.gpu
L0 = 6.283185307179586


def blob(forest, sigma=0.35):
    ks, _ = keys(forest)
    out = {}
    for (L, bx, by, bz) in ks:
        dx = L0 / (2 ** L * Bs)
        ax = [(b * Bs + np.arange(Bs)) * dx for b in (bx, by, bz)]
        Z, Y, X = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")
        r2 = (X - 0.5 * L0) ** 2 + (Y - 0.5 * L0) ** 2 + (Z - 0.5 * L0) ** 2
        f = 1.0 + 4.0 * np.exp(-r2 / (2.0 * sigma ** 2))
        out[(L, bx, by, bz)] = np.stack([f, 0.5 * f, -f, 2.0 * f])
    return out


def test_compression_sweep_matches_oracle():
    J = 3
    w = O.setup_wavelet("CDF40")
    p = tg_params(Bs=Bs, J=J, wavelet_g=w.g_default)
    p.wavelet = "CDF40"
    po = orc_params(p)
    uniform = Forest.uniform(3, J, Jmax=J)
    exact = blob(uniform)
    ks_u, hvy_u = keys(uniform)
    host0 = np.zeros((uniform.n_blocks, 4, Bs + 2 * p.g, Bs + 2 * p.g, Bs + 2 * p.g))
    g = p.g
    for key, h in zip(ks_u, hvy_u):
        host0[h - 1][:, g:g + Bs, g:g + Bs, g:g + Bs] = exact[key]
    norm = max(np.abs(v[0]).max() for v in exact.values())
    curve = []
    for eps in (1.0e-1, 1.0e-2, 1.0e-3, 1.0e-5):
        sol = WabbitGPU(p, max_blocks=uniform.n_blocks)
        sol.setup_wavelet("CDF40")
        sol.set_forest(uniform)
        sol.upload(host0)
        forest, data = uniform, dict(exact)
        for sweep in range(J):                                   # adapt_tree until nothing changes any more
            new, n0, n1 = sol.adapt_tree(forest, eps=eps, Jmin=1)
            data, st = orc_adapt(w, po, forest, data, eps, 1)
            assert n1 == len(data) and sorted(keys(new)[0]) == sorted(data)
            got = gpu_data(sol, new, po)
            assert all(np.array_equal(got[k], data[k]) for k in data), (eps, sweep)
            forest = new
            if n1 == n0:
                break
        nb = forest.n_blocks
        while not (forest.n_blocks == uniform.n_blocks):         # refineToEquidistant_tree
            data = orc_refine(w, po, forest, data)
            forest = sol.refine_tree(forest)
            got = gpu_data(sol, forest, po)
            assert sorted(got) == sorted(data) and all(np.array_equal(got[k], data[k]) for k in data), eps
        err = max(np.abs(got[k][0] - exact[k][0]).max() for k in exact) / norm
        curve.append((eps, nb, err))
        sol.close()
    # compression and error curves behave as the reference's test expects: fewer blocks and larger error for larger thresholds,
    # error of the order of the threshold (interpolating wavelets, Linfty normalisation)
    nbs, errs = [c[1] for c in curve], [c[2] for c in curve]
    assert nbs == sorted(nbs) and nbs[0] < nbs[-1] <= uniform.n_blocks
    assert all(a >= b for a, b in zip(errs, errs[1:]))
    assert all(e <= 10.0 * eps for eps, _, e in curve), curve
