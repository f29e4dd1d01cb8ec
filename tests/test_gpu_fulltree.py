"""adapt_tree with the full wavelet transformation for lifted wavelets (wabbit_b200/fulltree.py) against the oracle's restatement of
wavelet_decompose_full_tree / coarseningIndicator_tree / the grid decision / wavelet_reconstruct_full_tree_CEoptimized
(oracle/fulltree.py): coefficients of every block of the full tree (leaves and mothers), details and flags bit for bit."""
import numpy as np
import pytest

import fulltree as OFT
import oracle as O
from wabbit_b200 import Forest, WabbitGPU
from wabbit_b200.fulltree import FullTree, WD
from wabbit_b200.solver import HVY_BLOCK

from util import graded_blocks, orc_grid, orc_params, tg_params

pytestmark = pytest.mark.gpu


def _setup(wavelet, Bs, seed, Jmax=3, disc="FD_4th_central", noise=0.05):
    lv, ix = graded_blocks(3, 1, Jmax, seed, 0.3)
    forest = Forest.from_blocks(3, Jmax, lv, ix, max_blocks=2 * len(lv) + 16)
    w = O.setup_wavelet(wavelet)
    p = tg_params(Bs=Bs, J=Jmax, wavelet_g=w.g_default, discretization=disc)
    p.wavelet = wavelet
    grid, po = orc_grid(forest), orc_params(p)
    sol = WabbitGPU(p, max_blocks=forest.max_blocks)
    sol.setup_wavelet(wavelet)
    sol.set_forest(forest)
    u = O.alloc(grid, po)
    O.inicond_taylor_green(grid, po, u)
    # small scales in the lower half (quarter for unlifted wavelets) of the domain in x only: part of the tree is significant, part is not
    amp = np.where(grid.ixyz[:, 0] * (2 if w.lifted else 4) < 2 ** grid.level, noise, 1.0e-7)
    u += amp[:, None, None, None, None] * np.random.default_rng(seed).standard_normal(u.shape)
    g = po.g
    u[:, :, :g] = u[:, :, -g:] = 0.0                      # ghost nodes carry nothing: every value the transform reads is synchronised
    u[:, :, :, :g] = u[:, :, :, -g:] = 0.0
    u[:, :, :, :, :g] = u[:, :, :, :, -g:] = 0.0
    host = np.zeros(sol.host_shape())
    host[:grid.n] = u
    sol.upload(host, hvy_ids=np.arange(1, grid.n + 1, dtype=np.int32))
    H = {"FD_2nd_central": 1, "FD_4th_central": 2, "FD_6th_central": 3}[disc]
    return w, p, po, forest, grid, sol, u, H


@pytest.mark.parametrize("wavelet,Bs", [("CDF44", 16), ("CDF44", 18), ("CDF42", 16), ("CDF62", 16), ("CDF22", 16), ("CDF44", 22), ("CDF42", 26)])
def test_full_tree_decomposition_and_flags(wavelet, Bs):
    w, p, po, forest, grid, sol, u, H = _setup(wavelet, Bs, seed=3)
    ot = OFT.decompose_full_tree(po, w, grid, u, Jmin=1, fd_half_width=H)
    norm = O.norm_linfty_tree(po, u)
    ost = OFT.threshold_full_tree(ot, 0.01, norm=norm, level_ref=forest.Jmax)
    ft = FullTree(sol, forest, Jmin=1)
    assert ft.leaf_first == ot.leaf_first and set(ft.slot) == set(ot.blk) and len(ft.slot) > len(ft.leaf)
    gnorm = sol.componentWiseNorm_tree((HVY_BLOCK, 0))
    assert np.array_equal(gnorm, norm)
    st = ft.decompose(eps=0.01, norm=gnorm)
    slot, leaf = ft.slot, ft.leaf
    keys = sorted(slot, key=lambda k: slot[k])
    ids = np.array([slot[k] for k in keys], dtype=np.int32)
    wd = np.zeros(sol.host_shape())
    sol.download(wd, WD[0], WD[1], hvy_ids=ids, g_sync=0)
    uu = np.zeros(sol.host_shape())
    sol.download(uu, HVY_BLOCK, 0, hvy_ids=ids, g_sync=0)
    I = (slice(None),) + O.interior(po)
    for k in keys:
        assert np.array_equal(wd[slot[k] - 1][I], ot.blk[k][I]), k              # decomposed values (hvy_block of the reference)
        assert np.array_equal(uu[slot[k] - 1][I], ot.tmp[k][I]), k              # original / assembled values (hvy_tmp)
    assert st == ost
    assert 0 < sum(1 for v in st.values() if v == -1) < len(st)
    sol.close()


@pytest.mark.parametrize("wavelet,Bs,indicator,sz", [("CDF44", 16, "threshold-state-vector", False), ("CDF44", 18, "threshold-state-vector", False),
                                                      ("CDF42", 16, "threshold-state-vector", False), ("CDF44", 16, "everywhere", False),
                                                      ("CDF62", 20, "threshold-state-vector", False), ("CDF22", 16, "everywhere", False),
                                                      ("CDF44", 16, "threshold-state-vector", True), ("CDF42", 18, "threshold-state-vector", True),
                                                      ("CDF40", 16, "threshold-state-vector", False), ("CDF60", 18, "threshold-state-vector", False),
                                                      ("CDF44", 22, "threshold-state-vector", True), ("CDF44", 26, "threshold-state-vector", False),
                                                      # Bs < Nrecon (adapt_tree.f90:771-803): the same-level neighbours of interface blocks are reconstructed too
                                                      ("CDF44", 14, "threshold-state-vector", False), ("CDF62", 16, "threshold-state-vector", True)])
def test_adapt_tree_lifted(wavelet, Bs, indicator, sz):
    """the whole adapt_tree with the full wavelet transformation (decomposition of the full tree, indicator, grid decision; for lifted
    wavelets coarse extension on the lasting interfaces and CE-optimised reconstruction; pruning): same new grid as the oracle, data bit
    for bit; a second adapt_tree changes nothing (the reference's invertibility criterion, unit_test_waveletDecomposition_invertibility.f90)"""
    w, p, po, forest, grid, sol, u, H = _setup(wavelet, Bs, seed=5)
    norm = O.norm_linfty_tree(po, u)
    eps = 0.01
    og, od, oi = OFT.adapt_tree(po, w, grid, u, eps, Jmin=1, norm=norm, level_ref=forest.Jmax, indicator=indicator, fd_half_width=H,
                                use_security_zone=sz)
    ft = FullTree(sol, forest, Jmin=1)
    new, info = ft.adapt(eps=eps, norm=sol.componentWiseNorm_tree((HVY_BLOCK, 0)), indicator=indicator, use_security_zone=sz)
    assert info["leaf_first"] == oi["leaf_first"] and info["leaf_only"] == oi["leaf_only"]
    if not sz:
        assert info["status0"] == oi["status0"]
    assert {k: v == -1 for k, v in info["status"].items()} == {k: v == -1 for k, v in oi["status"].items()}
    assert info["marked"] == oi["marked"] and (len(info["marked"]) > 0) == bool(w.lifted)
    hvy, lvl, ixyz, _ = new.active(0)
    okey = {(int(og.level[b]),) + tuple(int(v) for v in og.ixyz[b]): b for b in range(og.n)}
    keys = [(int(l), int(x[0]), int(x[1]), int(x[2])) for l, x in zip(lvl, ixyz)]
    assert sorted(keys) == sorted(okey) and 8 <= new.n_blocks < forest.n_blocks
    got = np.zeros(sol.host_shape())
    sol.download(got, g_sync=0)
    I = (slice(None),) + O.interior(po)
    for h, k in zip(hvy, keys):
        assert np.array_equal(got[h - 1][I], od[okey[k]][I]), k
    # adapt(adapt(u)) = adapt(u): with eps = 0 nothing is coarsened any more and the filtered interfaces are a fixed point
    ft2 = FullTree(sol, new, Jmin=1)
    new2, info2 = ft2.adapt(eps=0.0, norm=None, use_security_zone=sz)
    assert new2.n_blocks == new.n_blocks
    got2 = np.zeros(sol.host_shape())
    sol.download(got2, g_sync=0)
    a, b = got2[:new.n_blocks][(slice(None),) + I], got[:new.n_blocks][(slice(None),) + I]
    assert abs(np.sqrt((a ** 2).sum()) / np.sqrt((b ** 2).sum()) - 1.0) <= 1.0e-14 and np.abs(a - b).max() <= 1.0e-13
    sol.close()


@pytest.mark.parametrize("wavelet,Bs,eps_norm", [("CDF44", 16, "L2"), ("CDF42", 18, "L1"), ("CDF44", 16, "H1")])
def test_adapt_tree_lifted_security_zone_other_norms(wavelet, Bs, eps_norm):
    """the security zone with eps_norm = L1 / L2 / H1: the strips' coefficients are renormalised as threshold_block does (wavelet_renorm_block)
    before they are compared with eps * norm (wgpu_patch_details_norm) -- same grid decision and data as the oracle"""
    w, p, po, forest, grid, sol, u, H = _setup(wavelet, Bs, seed=7)
    norm = sol.componentWiseNorm_tree((HVY_BLOCK, 0), eps_norm)
    eps = 0.02
    og, od, oi = OFT.adapt_tree(po, w, grid, u, eps, Jmin=1, norm=norm, eps_norm=eps_norm, level_ref=forest.Jmax, fd_half_width=H, use_security_zone=True)
    _, _, oi0 = OFT.adapt_tree(po, w, grid, u, eps, Jmin=1, norm=norm, eps_norm=eps_norm, level_ref=forest.Jmax, fd_half_width=H, use_security_zone=False)
    ft = FullTree(sol, forest, Jmin=1)
    new, info = ft.adapt(eps=eps, norm=norm, eps_norm=eps_norm, use_security_zone=True)
    assert {k: v == -1 for k, v in info["status"].items()} == {k: v == -1 for k, v in oi["status"].items()}
    hvy, lvl, ixyz, _ = new.active(0)
    okey = {(int(og.level[b]),) + tuple(int(v) for v in og.ixyz[b]): b for b in range(og.n)}
    keys = [(int(l), int(x[0]), int(x[1]), int(x[2])) for l, x in zip(lvl, ixyz)]
    assert sorted(keys) == sorted(okey) and 8 <= new.n_blocks < forest.n_blocks
    got = np.zeros(sol.host_shape())
    sol.download(got, g_sync=0)
    I = (slice(None),) + O.interior(po)
    for h, k in zip(hvy, keys):
        assert np.array_equal(got[h - 1][I], od[okey[k]][I]), k
    print(eps_norm, "blocks", forest.n_blocks, "->", new.n_blocks, "kept by the security zone:", oi0["n_deleted"] - oi["n_deleted"])
    sol.close()


@pytest.mark.parametrize("eps_norm,eps", [("Linfty", 1.0e-3), ("L2", 1.0e-3)])
def test_security_zone_keeps_the_neighbour_of_a_significant_strip(eps_norm, eps):
    """addSecurityZone_CE_tree (the reference's default for lifted wavelets): a narrow bump 6 points inside a block, next to a face, is invisible
    to the neighbour's own coefficients (its filters reach 3 points into the ghost nodes) but lies inside the Nwc = 8 deep strip of the
    block that holds it -- the neighbour would be coarsened without the security zone and is kept with it.  Oracle and device agree on
    both grids and on the data."""
    wavelet, Bs, J = "CDF44", 16, 2
    w = O.setup_wavelet(wavelet)
    forest = Forest.uniform(3, J, Jmax=J, max_blocks=100)
    p = tg_params(Bs=Bs, J=J, wavelet_g=w.g_default)
    p.wavelet = wavelet
    grid, po = orc_grid(forest), orc_params(p)
    g = po.g
    u = O.alloc(grid, po)
    h = 2.0 * np.pi / (2 ** J * Bs)
    c = np.array([(1 * Bs + 9) * h, (1 * Bs + 8) * h, (1 * Bs + 8) * h])       # block (1,1,1), 6 points from its +x face
    for b in range(grid.n):
        ax = [(int(grid.ixyz[b, a]) * Bs + np.arange(Bs)) * h for a in range(3)]
        Z, Y, X = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")
        bump = np.exp(-((X - c[0]) ** 2 + (Y - c[1]) ** 2 + (Z - c[2]) ** 2) / (2.0 * (0.8 * h) ** 2))
        u[b, :, g:-g, g:-g, g:-g] = 1.0 + np.stack([bump, 0.5 * bump, -bump, 2.0 * bump])
    res = {}
    for sz in (False, True):
        sol = WabbitGPU(p, max_blocks=100)
        sol.setup_wavelet(wavelet)
        sol.set_forest(forest)
        host = np.zeros(sol.host_shape())
        host[:grid.n] = u
        sol.upload(host, hvy_ids=np.arange(1, grid.n + 1, dtype=np.int32))
        norm = sol.componentWiseNorm_tree((HVY_BLOCK, 0), eps_norm)
        if eps_norm == "Linfty":
            assert np.array_equal(norm, O.norm_linfty_tree(po, u))
        og, od, oi = OFT.adapt_tree(po, w, grid, u, eps, Jmin=1, norm=norm, eps_norm=eps_norm, level_ref=J, fd_half_width=2, use_security_zone=sz)
        new, info = FullTree(sol, forest, Jmin=1).adapt(eps=eps, norm=norm, eps_norm=eps_norm, use_security_zone=sz)
        hvy, lvl, ixyz, _ = new.active(0)
        okey = {(int(og.level[b]),) + tuple(int(v) for v in og.ixyz[b]): b for b in range(og.n)}
        keys = [(int(l), int(x[0]), int(x[1]), int(x[2])) for l, x in zip(lvl, ixyz)]
        assert sorted(keys) == sorted(okey)
        got = np.zeros(sol.host_shape())
        sol.download(got, g_sync=0)
        I = (slice(None),) + O.interior(po)
        assert all(np.array_equal(got[hh - 1][I], od[okey[k]][I]) for hh, k in zip(hvy, keys))
        res[sz] = new.n_blocks
        sol.close()
    print(eps_norm, eps, res)
    assert res[False] < res[True] <= grid.n, res
