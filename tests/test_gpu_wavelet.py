"""GPU parity of the wavelet side through the C ABI: decomposition, reconstruction, detail norms / refinement flags and
the component-wise norm against the oracle.  The FWT -> renorm -> max|wc| -> compare chain is computed without FMA in
the reference's term order, so decomposed fields, details and flags must be BIT-EXACT; reconstruction likewise matches
the oracle's generic branch exactly."""
import numpy as np
import pytest

import oracle as O
from util import orc_grid, orc_params
from wabbit_b200 import HVY_BLOCK, HVY_TMP, HVY_WORK, Forest, Params, WabbitGPU

pytestmark = pytest.mark.gpu


def make(name, Bs, J, nc=4, seed=1):
    w = O.setup_wavelet(name)
    p = Params(dim=3, domain=(1.0, 1.0, 1.0), Bs=(Bs,) * 3, wavelet=name, g=w.g_default, g_rhs=2, n_eqn=nc, Jmax=J,
               discretization="FD_4th_central").finalize()
    forest = Forest.uniform(3, J)
    sol = WabbitGPU(p, max_blocks=forest.n_blocks)
    sol.set_forest(forest)
    g, _ = sol.setup_wavelet(name)
    assert g == w.g_default
    grid = orc_grid(forest)
    po = orc_params(p)
    return w, p, po, forest, sol, grid


def smooth_field(grid, po, rng, nc):
    """superposition of a few Fourier modes + small noise: details span many orders of magnitude"""
    u = O.alloc(grid, po, nc)
    g = po.g
    for b in range(grid.n):
        x0, dx = grid.spacing_origin(po, b)
        ax = [(np.arange(po.Bs[d] + 2 * g) - g) * dx[d] + x0[d] for d in range(3)]
        Z, Y, X = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")
        for c in range(nc):
            u[b, c] = np.sin(2 * np.pi * (c + 1) * X) * np.cos(2 * np.pi * Y) + 0.3 * np.cos(4 * np.pi * Z + c)
    u += 1e-6 * rng.random(u.shape)
    return u


def interior(p, a):
    g = p.g
    return a[:, :, g:g + p.Bs[2], g:g + p.Bs[1], g:g + p.Bs[0]]


@pytest.mark.parametrize("name,Bs", [("CDF44", 16), ("CDF44", 18), ("CDF40", 16), ("CDF42", 20), ("CDF22", 16), ("CDF62", 16), ("CDF20", 18)])
def test_decomposition_bit_exact(name, Bs):
    w, p, po, forest, sol, grid = make(name, Bs, 2)
    rng = np.random.default_rng(3)
    u = rng.random(sol.host_shape())
    sol.upload(u)
    sol.waveletDecomposition_tree()
    wd = np.zeros_like(u)
    sol.download(wd, HVY_TMP, g_sync=0)
    O.sync_ghosts_same_level(grid, po, u, po.g, po.g)
    ref = np.zeros_like(u)
    O.fwt_tree(w, po, u, ref)
    assert np.array_equal(interior(p, wd), interior(p, ref))
    # reconstruction: back to the input within round-off, and exactly the oracle's reconstruction of the same coefficients
    sol.waveletReconstruction_tree(src=(HVY_TMP, 0), dst=(HVY_WORK, 2))
    r = np.zeros_like(u)
    sol.download(r, HVY_WORK, 2, g_sync=0)
    O.sync_ghosts_same_level(grid, po, ref, po.g, po.g)
    rr = np.zeros_like(u)
    O.iwt_tree(w, po, ref, rr)
    assert np.array_equal(interior(p, r), interior(p, rr))
    err = np.sqrt(((interior(p, r) - interior(p, u)) ** 2).sum() / (interior(p, u) ** 2).sum())
    assert err <= 1e-14      # unit_test_waveletDecomposition.f90 pass criterion
    sol.close()


@pytest.mark.parametrize("eps_norm", ["Linfty", "L2", "L1", "H1"])
def test_threshold_flags_bit_exact(eps_norm):
    w, p, po, forest, sol, grid = make("CDF44", 16, 2)
    rng = np.random.default_rng(5)
    u = smooth_field(grid, po, rng, 4)
    sol.upload(u)
    norm_gpu = sol.componentWiseNorm_tree()
    norm_ref = O.norm_linfty_tree(po, u)
    assert np.array_equal(norm_gpu, norm_ref)
    sol.waveletDecomposition_tree()
    O.sync_ghosts_same_level(grid, po, u, po.g, po.g)
    ref = np.zeros_like(u)
    O.fwt_tree(w, po, u, ref)
    lvl = grid.level
    n_keep_total = 0
    for eps in (1e-8, 1e-5, 1e-3, 3e-2, 1.0):
        for tc in ([1, 1, 1, 1], [1, 0, 2, 2]):
            st, det = sol.threshold_tree(eps=eps, norm=norm_gpu, eps_norm=eps_norm, thresh_comp=tc, level_ref=2, want_detail=True)
            st_ref, det_ref = O.threshold_tree(po, ref, lvl, eps, norm_ref, eps_norm, tc, level_ref=2)
            assert np.array_equal(det, det_ref)
            assert np.array_equal(st, st_ref)
            n_keep_total += int((st == 0).sum())
    assert 0 < n_keep_total
    sol.close()


def test_full_size_wavelet_properties():
    """BASELINE config 5 shape (4096 blocks, Bs=16, CDF44, nc=4): IWT(FWT(u)) = u, a constant field has zero details
    (flags all -1), flags of random data all 0."""
    w, p, po, forest, sol, grid = make("CDF44", 16, 4)
    rng = np.random.default_rng(9)
    u = rng.random(sol.host_shape())
    sol.upload(u)
    sol.waveletDecomposition_tree()
    st = sol.threshold_tree(eps=1e-3)
    assert (st == 0).all()
    sol.waveletReconstruction_tree(src=(HVY_TMP, 0), dst=(HVY_WORK, 2))
    r = np.zeros_like(u)
    sol.download(r, HVY_WORK, 2, g_sync=0)
    a, b = interior(p, r), interior(p, u)
    assert np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()) <= 1e-14
    const = np.zeros_like(u)
    const[:, 0], const[:, 1], const[:, 2], const[:, 3] = 1.0, -2.0, 0.5, 3.0
    sol.upload(const)
    sol.waveletDecomposition_tree()
    st, det = sol.threshold_tree(eps=1e-12, want_detail=True)
    assert (st == -1).all() and np.abs(det).max() <= 1e-15
    sol.close()


@pytest.mark.parametrize("name", ["CDF20", "CDF22", "CDF40", "CDF42", "CDF44", "CDF60", "CDF62"])
@pytest.mark.parametrize("Bs", [16, 24])
def test_fast_and_generic_kernels_agree(name, Bs):
    """The compile-time specialised transform (wavelet_fast_kernel) and the generic, table-driven one give identical bits,
    forward and inverse."""
    import os
    w, p, po, forest, sol, grid = make(name, Bs, 1)
    rng = np.random.default_rng(11)
    u = rng.standard_normal(sol.host_shape())
    sol.upload(u)
    out = {}
    for mode in ("fast", "generic"):
        if mode == "generic":
            os.environ["WGPU_WAVELET_GENERIC"] = "1"
        try:
            sol.waveletDecomposition_tree()
            a = np.zeros_like(u)
            sol.download(a, HVY_TMP, g_sync=0)
            sol.waveletReconstruction_tree(src=(HVY_TMP, 0), dst=(HVY_WORK, 2))
            b = np.zeros_like(u)
            sol.download(b, HVY_WORK, 2, g_sync=0)
        finally:
            os.environ.pop("WGPU_WAVELET_GENERIC", None)
        out[mode] = (a, b)
    assert np.array_equal(out["fast"][0], out["generic"][0])
    assert np.array_equal(out["fast"][1], out["generic"][1])
    sol.close()


def test_componentwise_norms_l1_l2_on_graded_grid():
    """componentWiseNorm_tree L1 / L2 (volume-weighted sums over leaf interiors) against a NumPy restatement of
    componentWiseNorm_tree.f90:150-197, 283-290; Linfty bit-exact."""
    from util import graded_blocks
    from wabbit_b200 import Forest, WabbitGPU
    lv, ix = graded_blocks(3, 1, 3, seed=21)
    forest = Forest.from_blocks(3, 3, lv, ix)
    w = O.setup_wavelet("CDF40")
    from util import tg_params, orc_grid, orc_params
    p = tg_params(Bs=16, J=3, wavelet_g=3)
    p.wavelet = "CDF40"
    po, grid = orc_params(p), orc_grid(forest)
    sol = WabbitGPU(p, max_blocks=forest.n_blocks)
    sol.setup_wavelet("CDF40")
    sol.set_forest(forest)
    u = np.random.default_rng(2).standard_normal(sol.host_shape())
    sol.upload(u)
    I = O.interior(po)
    l1 = np.zeros(4)
    l2 = np.zeros(4)
    for b in range(grid.n):
        _, dx = grid.spacing_origin(po, b)
        dv = dx[0] * dx[1] * dx[2]
        for c in range(4):
            blk = u[b, c][I]
            l1[c] += dv * np.abs(blk).sum()
            l2[c] += dv * (blk ** 2).sum()
    l2 = np.sqrt(l2)
    assert np.allclose(sol.componentWiseNorm_tree(norm="L1"), l1, rtol=1e-12, atol=0)
    assert np.allclose(sol.componentWiseNorm_tree(norm="L2"), l2, rtol=1e-12, atol=0)
    assert np.array_equal(sol.componentWiseNorm_tree(norm="Linfty"), O.norm_linfty_tree(po, u[:grid.n]))
    a = sol.componentWiseNorm_tree(norm="L2")
    assert np.array_equal(a, sol.componentWiseNorm_tree(norm="L2"))      # deterministic
    sol.close()
