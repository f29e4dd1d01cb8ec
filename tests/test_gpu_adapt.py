"""GPU parity of the level-crossing heavy-data kernels (wabbit_b200/csrc/jump.cu) against the oracle:
download with a fully synchronised ghost shell on graded grids (all 26 relations; copy, decimation, prediction), refineBlock,
sync_D2M / executeCoarsening, and the reference's unit-test property Coarsen(Refine(u)) = u (unit_test_refineCoarsen.f90:129)."""
import numpy as np
import pytest

import oracle as O
from wabbit_b200 import Forest, WabbitGPU
from wabbit_b200.solver import HVY_BLOCK, HVY_WORK

from util import graded_blocks, orc_grid, orc_params, relerr, tg_params

pytestmark = pytest.mark.gpu


def _case(wavelet, Bs, forest, seed, max_blocks=None):
    w = O.setup_wavelet(wavelet)
    p = tg_params(Bs=Bs, J=forest.Jmax, wavelet_g=w.g_default)
    p.wavelet = wavelet
    grid = orc_grid(forest)
    po = orc_params(p)
    sol = WabbitGPU(p, max_blocks=max_blocks or forest.n_blocks)
    sol.setup_wavelet(wavelet)
    sol.set_forest(forest)
    rng = np.random.default_rng(seed)
    u = np.zeros(sol.host_shape())
    u[:grid.n] = rng.standard_normal((grid.n,) + u.shape[1:])
    return w, p, po, grid, sol, u


@pytest.mark.parametrize("wavelet,Bs,ignore_filter", [("CDF40", 16, True), ("CDF44", 16, True), ("CDF20", 16, True), ("CDF62", 20, True),
                                                      ("CDF44", 16, False), ("CDF42", 18, False), ("CDF22", 16, False), ("CDF62", 20, False),
                                                      ("CDF44", 24, False), ("CDF44", 22, False), ("CDF44", 26, False), ("CDF42", 32, False),
                                                      ("CDF40", 28, True)])
def test_download_with_ghosts_on_graded_grid(wavelet, Bs, ignore_filter):
    """ignore_filter=False is sync_ghosts_tree's default: with a lifted wavelet the restriction goes through the HD filter
    (restrict_copy_at_CE); compared at the full depth g (at smaller depths the reference's filter reads ghost nodes it did not fill)."""
    lv, ix = graded_blocks(3, 1, 3, seed=7)
    forest = Forest.from_blocks(3, 3, lv, ix)
    w, p, po, grid, sol, u = _case(wavelet, Bs, forest, seed=1)
    nbr = forest.neighbors(0)[:, :grid.n]
    sol.set_ghost_filter(ignore_filter)
    sol.upload(u)
    for gs in ((p.g, max(p.g_rhs, w.X // 2)) if ignore_filter else (p.g,)):   # minimum sync depth X/2 (ini_file_to_params.f90:467-468)
        got = u.copy()
        sol.download(got, g_sync=gs)
        ref = u.copy()
        O.sync_ghosts_leaf(grid, po, ref, nbr, gs, gs, w.X, bool(w.lifted), ignore_filter=ignore_filter, w=w)
        assert np.array_equal(got, ref), gs
        if not ignore_filter:
            plain = u.copy()
            O.sync_ghosts_leaf(grid, po, plain, nbr, gs, gs, w.X, True)
            assert not np.array_equal(plain, ref)      # the filter does change ghost values on this grid
    sol.close()


def _refine_oracle(w, po, grid, u, nbr, ignore_filter=True):
    """sync_ghosts_tree + refineBlock per block: dict (level, ix, iy, iz) -> daughter interior [nc, Bs, Bs, Bs]"""
    ref = u.copy()
    if nbr is None:
        O.sync_ghosts_same_level(grid, po, ref, po.g, po.g)
    else:
        O.sync_ghosts_leaf(grid, po, ref, nbr, po.g, po.g, w.X, bool(w.lifted), ignore_filter=ignore_filter, w=w)
    I = O.interior(po)
    out = {}
    for b in range(grid.n):
        d = O.refine_block(w.X, po, ref[b])
        L, (x, y, z) = int(grid.level[b]), (int(v) for v in grid.ixyz[b])
        for k in range(8):
            q = ((k >> 1) & 1, k & 1, (k >> 2) & 1)
            out[(L + 1, 2 * x + q[0], 2 * y + q[1], 2 * z + q[2])] = d[k][(slice(None),) + I]
    return out


@pytest.mark.parametrize("wavelet,Bs", [("CDF40", 16), ("CDF44", 16), ("CDF20", 18), ("CDF62", 20), ("CDF44", 26), ("CDF40", 32)])
def test_refine_everywhere_uniform(wavelet, Bs):
    forest = Forest.uniform(3, 1, Jmax=2)
    w, p, po, grid, sol, u = _case(wavelet, Bs, forest, seed=2, max_blocks=64)
    sol.upload(u)
    expect = _refine_oracle(w, po, grid, u[:grid.n], None)
    new = sol.refine_tree(forest)
    assert new.n_blocks == 64
    got = np.zeros(sol.host_shape())
    sol.download(got, g_sync=0)
    hvy, lvl, ixyz, _ = new.active(0)
    I = (slice(None),) + O.interior(po)
    for h, l, (x, y, z) in zip(hvy, lvl, ixyz):
        assert np.array_equal(got[h - 1][I], expect[(int(l), int(x), int(y), int(z))])
    sol.close()


@pytest.mark.parametrize("wavelet,ignore_filter", [("CDF40", True), ("CDF44", True), ("CDF44", False)])
def test_refine_everywhere_graded_and_partial(wavelet, ignore_filter):
    # everywhere on a graded grid: mothers next to coarser blocks are interpolated from predicted ghost nodes, mothers next to
    # finer blocks from restricted ones (through the HD filter for a lifted wavelet, as the sync_ghosts_tree of main.f90:314)
    lv, ix = graded_blocks(3, 1, 2, seed=4)
    forest = Forest.from_blocks(3, 3, lv, ix)
    assert not forest.is_uniform
    w, p, po, grid, sol, u = _case(wavelet, 16, forest, seed=3, max_blocks=8 * forest.n_blocks)
    nbr = forest.neighbors(0)[:, :grid.n]
    sol.set_ghost_filter(ignore_filter)
    sol.upload(u)
    expect = _refine_oracle(w, po, grid, u[:grid.n], nbr, ignore_filter)
    new = sol.refine_tree(forest)
    got = np.zeros(sol.host_shape())
    sol.download(got, g_sync=0)
    hvy, lvl, ixyz, _ = new.active(0)
    I = (slice(None),) + O.interior(po)
    for h, l, (x, y, z) in zip(hvy, lvl, ixyz):
        assert np.array_equal(got[h - 1][I], expect[(int(l), int(x), int(y), int(z))])
    sol.close()

    # partial refinement of a uniform grid: blocks that stay keep their data (and move to their new slots)
    forest = Forest.uniform(3, 2, Jmax=3)
    w, p, po, grid, sol, u = _case("CDF44", 16, forest, seed=5, max_blocks=64 + 7 * 20)
    sol.upload(u)
    expect = _refine_oracle(w, po, grid, u[:grid.n], None)
    flags = np.zeros(grid.n, np.int32)
    flags[np.random.default_rng(0).choice(grid.n, 20, replace=False)] = 1
    new = sol.refine_tree(forest, flags)
    assert new.n_blocks == 64 + 7 * 20
    got = np.zeros(sol.host_shape())
    sol.download(got, g_sync=0)
    old = {(int(l), int(a), int(b), int(c)): k for k, (l, (a, b, c)) in enumerate(zip(grid.level, grid.ixyz))}
    hvy, lvl, ixyz, _ = new.active(0)
    I = (slice(None),) + O.interior(po)
    for h, l, (x, y, z) in zip(hvy, lvl, ixyz):
        key = (int(l), int(x), int(y), int(z))
        if key in old:
            assert np.array_equal(got[h - 1][I], u[old[key]][I])
        else:
            assert np.array_equal(got[h - 1][I], expect[key])
    sol.close()


@pytest.mark.parametrize("wavelet,Bs", [("CDF40", 16), ("CDF44", 16), ("CDF42", 18), ("CDF62", 20)])
def test_coarsen_of_refine_is_identity_and_matches_oracle(wavelet, Bs):
    forest = Forest.uniform(3, 1, Jmax=2)
    w, p, po, grid, sol, u = _case(wavelet, Bs, forest, seed=6, max_blocks=64)
    sol.upload(u)
    fine = sol.refine_tree(forest)
    fine_host = np.zeros(sol.host_shape())
    sol.download(fine_host, g_sync=p.g)
    sol.waveletDecomposition_tree(src=(HVY_BLOCK, 0), dst=(HVY_WORK, 2))
    status = np.full(fine.n_blocks, -1, np.int32)
    coarse = sol.executeCoarsening_tree(fine, status, decomposed=(HVY_WORK, 2))
    assert coarse.n_blocks == 8 and coarse.is_uniform
    got = np.zeros(sol.host_shape())
    sol.download(got, g_sync=0)
    hvy, lvl, ixyz, _ = coarse.active(0)
    I = (slice(None),) + O.interior(po)
    old = {(int(l), int(a), int(b), int(c)): k for k, (l, (a, b, c)) in enumerate(zip(grid.level, grid.ixyz))}
    # (1) oracle: decomposition of the refined field, scaling coefficients at even positions -> octants of the mothers
    fgrid = orc_grid(fine)
    wd = np.zeros_like(fine_host[:fgrid.n])
    O.fwt_tree(w, po, fine_host[:fgrid.n], wd)
    g, h = po.g, Bs // 2
    for hm, l, (x, y, z) in zip(hvy, lvl, ixyz):
        exp = np.zeros((4, Bs, Bs, Bs))
        for k, (fl, (fx, fy, fz)) in enumerate(zip(fgrid.level, fgrid.ixyz)):
            if (fx // 2, fy // 2, fz // 2) == (x, y, z):
                qx, qy, qz = fx % 2, fy % 2, fz % 2
                exp[:, qz * h:(qz + 1) * h, qy * h:(qy + 1) * h, qx * h:(qx + 1) * h] = wd[k][:, g:g + Bs:2, g:g + Bs:2, g:g + Bs:2]
        assert np.array_equal(got[hm - 1][I], exp)
        # (2) the reference's property test: Coarsen(Refine(u)) = u to 1e-14
        assert relerr(got[hm - 1][I], u[old[(int(l), int(x), int(y), int(z))]][I]) <= 1e-14
    sol.close()


def test_page_locked_host_arrays_take_the_direct_path_with_identical_results():
    """wgpu_upload / wgpu_download on a page-locked host array (kernels read / write it over PCIe, no staging) == staged path,
    on a uniform and on a graded grid, for a full and a partial ghost shell."""
    import torch
    for graded in (False, True):
        if graded:
            lv, ix = graded_blocks(3, 1, 3, seed=9)
            forest = Forest.from_blocks(3, 3, lv, ix)
        else:
            forest = Forest.uniform(3, 2, Jmax=3)
        w, p, po, grid, sol, u = _case("CDF44", 16, forest, seed=8)
        pinned = torch.empty(u.shape, dtype=torch.float64, pin_memory=True)
        pn = pinned.numpy()
        pn[:] = u
        sol.set_transfer_mode(False, False)        # zero-copy layout kernels (the copy engines: next test)
        n0 = sol.launch_count
        sol.upload(pn)
        assert sol.launch_count - n0 == 1          # one kernel, no staging chunks
        for gs in (0, 2, p.g):
            a = np.full_like(u, -7.0)
            sol.download(a, g_sync=gs)
            pn[:] = -7.0
            sol.download(pn, g_sync=gs)
            assert np.array_equal(a, pn), (graded, gs)
        sol.close()


@pytest.mark.parametrize("Bs,wavelet", [(16, "CDF44"), (18, "CDF40"), (22, "CDF42")])
def test_copy_engine_transfers_of_page_locked_host_arrays(Bs, wavelet):
    """default path of wgpu_upload / wgpu_download(g_sync=0) for page-locked 3-D arrays (wgpu_set_transfer_mode): plane spans by DMA + a
    layout kernel.  Interiors are exact; on a download the x ghost nodes between the interior rows hold the same-level neighbour's values;
    nothing outside the spans is touched."""
    import torch
    for graded in (False, True):
        if graded:
            lv, ix = graded_blocks(3, 1, 3, seed=9)
            forest = Forest.from_blocks(3, 3, lv, ix)
        else:
            forest = Forest.uniform(3, 2, Jmax=3)
        w, p, po, grid, sol, u = _case(wavelet, Bs, forest, seed=8)
        g, n = p.g, grid.n
        pinned = torch.empty(u.shape, dtype=torch.float64, pin_memory=True)
        pn = pinned.numpy()
        pn[:] = u
        sol.upload(pn)                                         # copy engines
        a = np.full_like(u, -7.0)
        sol.download(a, g_sync=0)                              # pageable: staged path
        I = (slice(0, n), slice(None)) + O.interior(po)
        assert np.array_equal(a[I], u[I]), graded
        pn[:] = -7.0
        sol.download(pn, g_sync=0)                             # copy engines
        assert np.array_equal(pn[I], u[I]), graded
        span = np.zeros(u.shape[2:], dtype=bool)               # [z, y, x]: the nodes a plane span covers
        flat = span[g:g + Bs].reshape(Bs, -1)
        nx = Bs + 2 * g
        flat[:, g * nx + g: (g + Bs - 1) * nx + g + Bs] = True
        assert (pn[:n][:, :, ~span] == -7.0).all(), graded
        if not graded:
            ref = u.copy()
            O.sync_ghosts_leaf(grid, po, ref, forest.neighbors(0)[:, :n], g, g, w.X, True)
            assert np.array_equal(pn[:n][:, :, span], ref[:n][:, :, span])
        sol.close()


@pytest.mark.parametrize("wavelet,Bs,ignore_filter", [("CDF40", 16, True), ("CDF44", 16, True), ("CDF22", 18, True), ("CDF62", 20, True),
                                                      ("CDF20", 24, True), ("CDF44", 16, False), ("CDF42", 20, False), ("CDF62", 16, False),
                                                      ("CDF44", 22, False), ("CDF44", 26, False), ("CDF40", 32, True), ("CDF42", 28, False)])
def test_wavelet_transform_on_graded_grid(wavelet, Bs, ignore_filter):
    """FWT / IWT on a leaf grid with level jumps: ghost values of all 26 relations come from the wavelet jump pool (restriction
    from finer -- through the HD filter unless ignore_filter --, prediction from coarser neighbours) and are exactly what the
    oracle's sync_ghosts_generic leaves, so the coefficients agree bit for bit with sync + waveletDecomposition_optimized_block
    on the host (the leaf decomposition of wavelet_decompose_full_tree, adapt_tree.f90:403-446)."""
    from wabbit_b200.solver import HVY_TMP
    lv, ix = graded_blocks(3, 1, 3, seed=12)
    forest = Forest.from_blocks(3, 3, lv, ix)
    w, p, po, grid, sol, u = _case(wavelet, Bs, forest, seed=13)
    nbr = forest.neighbors(0)[:, :grid.n]
    I = (slice(None), slice(None)) + O.interior(po)
    sol.set_ghost_filter(ignore_filter)
    sol.upload(u)
    sol.waveletDecomposition_tree((HVY_BLOCK, 0), (HVY_TMP, 0))
    wd = np.zeros_like(u)
    sol.download(wd, HVY_TMP, g_sync=0)
    ref = u[:grid.n].copy()
    O.sync_ghosts_leaf(grid, po, ref, nbr, po.g, po.g, w.X, bool(w.lifted), ignore_filter=ignore_filter, w=w)
    wd_ref = np.zeros_like(ref)
    O.fwt_tree(w, po, ref, wd_ref)
    assert np.array_equal(wd[:grid.n][I], wd_ref[I])
    # flags from the fused Linfty details == oracle's threshold_block
    st, det = sol.threshold_tree((HVY_TMP, 0), eps=0.5, want_detail=True)
    st_ref, det_ref = O.threshold_tree(po, wd_ref, grid.level, 0.5)
    assert np.array_equal(det, det_ref) and np.array_equal(st, st_ref)
    # inverse transform of the coefficient field, ghosts synchronised the same way
    sol.waveletReconstruction_tree(src=(HVY_TMP, 0), dst=(HVY_WORK, 2))
    r = np.zeros_like(u)
    sol.download(r, HVY_WORK, 2, g_sync=0)
    O.sync_ghosts_leaf(grid, po, wd_ref, nbr, po.g, po.g, w.X, bool(w.lifted), ignore_filter=ignore_filter, w=w)
    r_ref = np.zeros_like(wd_ref)
    O.iwt_tree(w, po, wd_ref, r_ref)
    assert np.array_equal(r[:grid.n][I], r_ref[I])
    sol.close()


@pytest.mark.parametrize("wavelet,Bs,disc", [("CDF44", 16, "FD_4th_central"), ("CDF42", 18, "FD_6th_central"), ("CDF40", 16, "FD_4th_central"),
                                             ("CDF62", 20, "FD_4th_central"), ("CDF22", 16, "FD_2nd_central")])
def test_coarse_extension_modify(wavelet, Bs, disc):
    """coarse_extension_modify on a graded grid: zeroed wavelet coefficients and copied scaling coefficients in the strips facing
    coarser neighbours, against the oracle's restatement of coarseExtensionManipulateWC_block / ...SC_block (pure copies: exact)."""
    from wabbit_b200.solver import HVY_TMP
    lv, ix = graded_blocks(3, 1, 3, seed=31)
    forest = Forest.from_blocks(3, 3, lv, ix)
    w = O.setup_wavelet(wavelet)
    p = tg_params(Bs=Bs, J=3, wavelet_g=w.g_default, discretization=disc)
    p.wavelet = wavelet
    grid, po = orc_grid(forest), orc_params(p)
    sol = WabbitGPU(p, max_blocks=forest.n_blocks)
    sol.setup_wavelet(wavelet)
    sol.set_forest(forest)
    rng = np.random.default_rng(3)
    orig = rng.standard_normal(sol.host_shape())
    wd = rng.standard_normal(sol.host_shape())
    nbr = forest.neighbors(0)
    H = {"FD_2nd_central": 1, "FD_4th_central": 2, "FD_6th_central": 3}[disc]
    I = (slice(None), slice(None)) + O.interior(po)
    for clear_wc, copy_sc in ((True, True), (True, False), (False, True)):
        sol.upload(orig)
        sol.upload(wd, HVY_TMP)
        sol.coarse_extension_modify((HVY_TMP, 0), (HVY_BLOCK, 0), clear_wc, copy_sc)
        got = np.zeros_like(wd)
        sol.download(got, HVY_TMP, g_sync=0)
        ref = wd.copy()
        n = O.coarse_extension_modify(grid, po, w, ref, orig, nbr, fd_half_width=H, clear_wc=clear_wc, copy_sc=copy_sc)
        assert n > 0
        assert np.array_equal(got[I], ref[I])
        if clear_wc or w.Nscr > 0:
            assert not np.array_equal(got[I], wd[I])
    sol.close()


@pytest.mark.parametrize("wavelet,Bs,disc", [("CDF44", 16, "FD_4th_central"), ("CDF42", 18, "FD_4th_central"), ("CDF62", 20, "FD_6th_central")])
def test_leaf_coarsening_indicator_lifted(wavelet, Bs, disc):
    """The leaf pass of adapt_tree for a lifted wavelet (wavelet_decompose_full_tree, iteration 0, LIB/MESH/adapt_tree.f90:403-470, and
    coarseningIndicator_tree): sync_TMP_from_all (restriction through the HD filter) -> waveletDecomposition_optimized_block ->
    coarse_extension_modify (zero WC / copy SC next to coarser neighbours) -> threshold_block.  Coefficients, details and refinement
    flags of every leaf equal the oracle's bit for bit."""
    from wabbit_b200.solver import HVY_TMP
    lv, ix = graded_blocks(3, 1, 3, seed=41)
    forest = Forest.from_blocks(3, 3, lv, ix)
    w = O.setup_wavelet(wavelet)
    p = tg_params(Bs=Bs, J=3, wavelet_g=w.g_default, discretization=disc)
    p.wavelet = wavelet
    grid, po = orc_grid(forest), orc_params(p)
    sol = WabbitGPU(p, max_blocks=forest.n_blocks)
    sol.setup_wavelet(wavelet)
    sol.set_forest(forest)
    nbr = forest.neighbors(0)[:, :grid.n]
    H = {"FD_2nd_central": 1, "FD_4th_central": 2, "FD_6th_central": 3}[disc]
    u = O.alloc(grid, po)
    O.inicond_taylor_green(grid, po, u)
    amp = np.where(grid.ixyz[:, 0] * 4 < 2 ** grid.level, 0.05, 1.0e-7)      # small scales in the lowest quarter in x only: mixed flags
    u += amp[:, None, None, None, None] * np.random.default_rng(4).standard_normal(u.shape)
    sol.upload(u)
    norm = sol.componentWiseNorm_tree((HVY_BLOCK, 0))
    sol.waveletDecomposition_tree((HVY_BLOCK, 0), (HVY_TMP, 0))
    sol.coarse_extension_modify((HVY_TMP, 0), (HVY_BLOCK, 0))
    st, det = sol.threshold_tree((HVY_TMP, 0), eps=0.01, norm=norm, want_detail=True, level_ref=3)
    wd = np.zeros_like(u)
    sol.download(wd, HVY_TMP, g_sync=0)
    ref = u.copy()
    O.sync_ghosts_leaf(grid, po, ref, nbr, po.g, po.g, w.X, True, ignore_filter=False, w=w)
    wd_ref = np.zeros_like(ref)
    O.fwt_tree(w, po, ref, wd_ref)
    n = O.coarse_extension_modify(grid, po, w, wd_ref, ref, nbr, fd_half_width=H)
    assert n > 0
    I = (slice(None), slice(None)) + O.interior(po)
    assert np.array_equal(wd[I], wd_ref[I])
    norm_ref = O.norm_linfty_tree(po, u)
    assert np.array_equal(norm, norm_ref)
    st_ref, det_ref = O.threshold_tree(po, wd_ref, grid.level, 0.01, norm=norm_ref, level_ref=3)
    assert np.array_equal(det, det_ref) and np.array_equal(st, st_ref)
    assert 0 < (st == -1).sum() < grid.n
    sol.close()
