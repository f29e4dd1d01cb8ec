"""Pin the oracle's restatement of adapt_tree with the full wavelet transformation (oracle/fulltree.py) with the properties the
reference's own unit tests check:
  * adapt(adapt(u)) = adapt(u) on a non-equidistant grid with random data, relative L2 change <= 1e-14
    (unit_test_waveletDecomposition_invertibility.f90; the first application does change the field at the coarse/fine interfaces),
    for both decomposition variants (leaf-first / level-wise) and both reconstruction variants (all leaves at once / level by level);
  * Coarsen(Refine(u)) = u (unit_test_refineCoarsen.f90:129, and the TESTING/wavelets/adaptive_CDFXY cases: --refine-everywhere followed
    by --coarsen-everywhere returns the input grid and field);
  * the grid decision leaves a graded grid and removes whole sister groups only."""
import numpy as np
import pytest

import fulltree as FT
import oracle as O

from util import graded_blocks


def _case(name, Bs, dim, seed, Jmax=3, nc=1):
    w = O.setup_wavelet(name)
    p = O.Params(dim=dim, Bs=(Bs, Bs, Bs if dim == 3 else 1), g=w.g_default, g_rhs=w.g_default, n_eqn=nc, Jmax=Jmax + 1)
    lv, ix = graded_blocks(dim, 1, Jmax, seed, 0.3)
    grid = O.Grid(level=lv.astype(np.int64), ixyz=ix.astype(np.int64), dim=dim)
    return w, p, grid


@pytest.mark.parametrize("name,Bs,dim", [("CDF44", 16, 2), ("CDF44", 16, 3), ("CDF44", 18, 3), ("CDF42", 16, 2), ("CDF22", 8, 2),
                                         ("CDF62", 18, 2), ("CDF44", 20, 2), ("CDF62", 16, 2), ("CDF44", 14, 2)])     # the last two: Bs < Nrecon
def test_adapt_of_adapt_is_adapt(name, Bs, dim):
    w, p, grid = _case(name, Bs, dim, seed=7, Jmax=3 if dim == 2 else 2)
    I = (slice(None), slice(None)) + O.interior(p)
    u = O.alloc(grid, p)
    u[I] = np.random.default_rng(0).random(u[I].shape)
    sz = Bs % 4 == 0                                          # with and without the security zone (the reference's test runs with it)
    g1, d1, i1 = FT.adapt_tree(p, w, grid, u, eps=1.0e-3, Jmin=1, use_security_zone=sz)
    assert g1.n == grid.n and len(i1["marked"]) > 0          # random data: nothing is coarsened, interfaces are filtered
    assert np.abs(d1[I] - u[I]).max() > 1.0e-2               # ... which changes the field
    unmarked = [b for b in range(g1.n) if (int(g1.level[b]),) + tuple(int(v) for v in g1.ixyz[b]) not in set(i1["marked"])]
    assert np.array_equal(d1[unmarked][I], u[unmarked][I])    # blocks away from interfaces keep their values exactly
    u1 = np.zeros_like(d1)
    u1[I] = d1[I]
    g2, d2, i2 = FT.adapt_tree(p, w, g1, u1, eps=1.0e-3, Jmin=1, use_security_zone=sz)
    n1, n2 = np.sqrt((d1[I] ** 2).sum()), np.sqrt((d2[I] ** 2).sum())
    assert abs(n2 / n1 - 1.0) <= 1.0e-14                     # the reference's criterion
    assert np.abs(d2[I] - d1[I]).max() <= 1.0e-14            # and pointwise


def _refine_everywhere(w, p, grid, u):
    nbr = O.neighbor_table168(grid, int(grid.level.max()) + 2)
    v = u.copy()
    O.sync_ghosts_leaf(grid, p, v, nbr, p.g, p.g, w.X, bool(w.lifted), ignore_filter=False, w=w)
    lv, ix, data = [], [], []
    nd = 2 ** grid.dim
    for b in range(grid.n):
        d = O.refine_block(w.X, p, v[b])
        L, (x, y, z) = int(grid.level[b]), (int(q) for q in grid.ixyz[b])
        for k in range(nd):
            q = ((k >> 1) & 1, k & 1, (k >> 2) & 1 if grid.dim == 3 else 0)
            lv.append(L + 1)
            ix.append((2 * x + q[0], 2 * y + q[1], 2 * z + q[2] if grid.dim == 3 else 0))
            data.append(d[k])
    order = sorted(range(len(lv)), key=lambda i: (lv[i],) + ix[i])
    g2 = O.Grid(level=np.array([lv[i] for i in order], dtype=np.int64), ixyz=np.array([ix[i] for i in order], dtype=np.int64), dim=grid.dim)
    return g2, np.stack([data[i] for i in order])


@pytest.mark.parametrize("name,Bs,dim", [("CDF44", 16, 2), ("CDF42", 16, 2), ("CDF44", 16, 3), ("CDF22", 12, 2)])
def test_coarsen_everywhere_of_refine_everywhere(name, Bs, dim):
    """On an equidistant grid the round trip is the identity for random data (<= 1e-13).  On a graded grid it returns the input GRID, and
    the input field up to the interpolation error at the level jumps: a block next to a finer one is filtered together with genuinely
    finer data there (the low-pass filter of a lifted wavelet is not a decimation), so only a smooth field comes back closely."""
    key = lambda g, b: (int(g.level[b]),) + tuple(int(v) for v in g.ixyz[b])
    for graded in (False, True):
        if graded:
            w, p, grid = _case(name, Bs, dim, seed=3, Jmax=3 if dim == 2 else 2)
        else:
            w, p, _ = _case(name, Bs, dim, seed=3)
            grid = O.uniform_grid(2, dim)
        I = (slice(None), slice(None)) + O.interior(p)
        u = O.alloc(grid, p)
        if graded:
            for b in range(grid.n):
                dx = 2.0 * np.pi / (2 ** int(grid.level[b]) * Bs)
                ax = [(int(grid.ixyz[b, a]) * Bs + np.arange(Bs)) * dx for a in range(dim)]
                if dim == 3:
                    Z, Y, X = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")
                    u[b][I[1:]] = np.sin(X) * np.cos(Y) * np.cos(Z)
                else:
                    Y, X = np.meshgrid(ax[1], ax[0], indexing="ij")
                    u[b][I[1:]] = (np.sin(X) * np.cos(Y))[None]
        else:
            u[I] = np.random.default_rng(1).random(u[I].shape)
        gf, uf = _refine_everywhere(w, p, grid, u)
        assert gf.n == grid.n * 2 ** dim
        uf0 = np.zeros_like(uf)
        uf0[I] = uf[I]
        gc, uc, info = FT.adapt_tree(p, w, gf, uf0, eps=0.0, Jmin=1, indicator="everywhere")
        assert sorted(key(gc, b) for b in range(gc.n)) == sorted(key(grid, b) for b in range(grid.n))
        pos = {key(grid, b): b for b in range(grid.n)}
        err = max(np.abs(uc[b][I[1:]] - u[pos[key(gc, b)]][I[1:]]).max() for b in range(gc.n))
        assert err <= ((1.0e-2 if w.X == 2 else 2.0e-3) if graded else 1.0e-13), (graded, err)


def test_decision_keeps_the_grid_graded_and_sister_groups_whole():
    w, p, grid = _case("CDF44", 16, 2, seed=11, Jmax=4)
    u = O.alloc(grid, p)
    t = FT.Tree(p, w, grid, u, Jmin=1)
    rng = np.random.default_rng(5)
    st0 = {k: (-1 if rng.random() < 0.7 else 0) for k in t.blk}
    st = FT.decide(t, st0, Jmin=1)
    gone = {k for k, v in st.items() if v == -1}
    assert gone and all(st0[k] == -1 for k in gone)
    for k in gone:
        assert all(s in gone for s in FT.children(FT.parent(k), 2))                      # completeness
        assert all(c in gone for c in FT.children(k, 2) if c in t.blk)                   # no orphan daughters
        assert k[0] > 1
    left = set(t.blk) - gone
    leaves = {k for k in left if not any(c in left for c in FT.children(k, 2))}
    for k in leaves:                                                                     # gradedness of the new leaf grid
        for d in FT.dirs(2):
            nk = FT.nbr_key(k, d, 2)
            if nk in left:
                continue
            ck = FT.parent(nk)
            assert ck in leaves, (k, d)                                                  # one level coarser at most


def test_security_zone_only_keeps_blocks():
    """addSecurityZone_CE_tree can only turn -1 into 0: the adapted grid with it contains every block of the adapted grid without it or its
    descendants, and a narrow feature next to a block face does make a difference (see tests/test_gpu_fulltree.py for the 3-D case)"""
    w, p, _ = _case("CDF44", 16, 2, seed=1)
    grid = O.uniform_grid(3, 2)
    I = (slice(None), slice(None)) + O.interior(p)
    u = O.alloc(grid, p)
    h = 1.0 / (8 * 16)
    for b in range(grid.n):
        ax = [(int(grid.ixyz[b, a]) * 16 + np.arange(16)) * h for a in range(2)]
        Y, X = np.meshgrid(ax[1], ax[0], indexing="ij")
        u[b][I[1:]] = 1.0 + np.exp(-((X - (3 * 16 + 9) * h) ** 2 + (Y - (3 * 16 + 8) * h) ** 2) / (2 * (0.8 * h) ** 2))[None]
    n = {}
    for sz in (False, True):
        g1, d1, i1 = FT.adapt_tree(p, w, grid, u, eps=1.0e-3, Jmin=1, use_security_zone=sz)
        n[sz] = g1.n
        if sz:
            assert all(v0 == -1 or i1["status0"][k] == 0 for k, v0 in st_prev.items() if v0 == 0)
        st_prev = i1["status0"]
    assert n[False] < n[True]
