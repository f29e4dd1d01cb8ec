"""Pin the oracle's level-jump ghost synchronisation with the reference's own unit test protocol (unit_test_Sync.f90:97-273):
two-level grid, linear integer-valued field, ghosts poisoned with -1, sync for every depth g_sync, every ghost value inside
g_sync must be exact and nothing outside may be touched.  Also cross-checks the product's host forest (libwabbit_host.so)
against the oracle's independent NumPy neighbour search."""
import numpy as np
import pytest

import oracle as O
from wabbit_b200 import Forest


def two_level_grid(dim, refine=((0, 0, 0),)):
    lv, ix = [], []
    for z in range(2 if dim == 3 else 1):
        for y in range(2):
            for x in range(2):
                if (x, y, z) in refine:
                    for c in range(2 ** dim):
                        lv.append(2)
                        ix.append((2 * x + (c & 1), 2 * y + ((c >> 1) & 1), 2 * z + ((c >> 2) & 1) if dim == 3 else 0))
                else:
                    lv.append(1)
                    ix.append((x, y, z))
    return O.Grid(level=np.array(lv, dtype=np.int64), ixyz=np.array(ix, dtype=np.int64), dim=dim)


def linear_field(grid, p, hvy):
    """integer-valued linear function of the position in units of the finest spacing (unit_test_fill_linearly)"""
    g = p.g
    Jfine = int(grid.level.max())
    for b in range(grid.n):
        J = int(grid.level[b])
        s = 2 ** (Jfine - J)
        ax = []
        for d in range(3):
            n = p.Bs[d] + 2 * g if d < grid.dim else 1
            off = (np.arange(n) - (g if d < grid.dim else 0) + int(grid.ixyz[b, d]) * p.Bs[d]) * s
            ax.append(off.astype(np.float64))
        Z, Y, X = np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")
        hvy[b, 0] = X + 3.0 * Y + 7.0 * Z
        if hvy.shape[1] > 1:
            hvy[b, 1] = 5.0 * X - 2.0 * Y + Z + 11.0


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("wavelet,Bs", [("CDF40", 16), ("CDF44", 16), ("CDF20", 12), ("CDF62", 20)])
def test_linear_field_sync_protocol(dim, wavelet, Bs):
    w = O.setup_wavelet(wavelet)
    g = w.g_default
    order = w.X
    g_rhs = max(w.X // 2, 1)
    p = O.Params(dim=dim, Bs=(Bs, Bs, Bs if dim == 3 else 1), g=g, g_rhs=g_rhs, n_eqn=2, Jmax=2)
    grid = two_level_grid(dim, refine=((0, 0, 0), (1, 1, 0)))
    nbr = O.neighbor_table168(grid, 2)
    interior = (slice(None), slice(None)) + O.interior(p)
    Jfine = 2
    for gs in range(g, g_rhs - 1, -1):
        expected = O.alloc(grid, p)
        linear_field(grid, p, expected)
        u = np.full_like(expected, -1.0)
        u[interior] = expected[interior]
        n = O.sync_ghosts_leaf(grid, p, u, nbr, gs, gs, order, bool(w.lifted))
        assert n > 0
        # points whose stencil crosses the periodic boundary are not comparable (the field jumps there)
        gmin = max(w.hd_hi, w.hr_hi + 1)
        bad = 0
        for b in range(grid.n):
            J = int(grid.level[b])
            s = 2 ** (Jfine - J)
            ntot = 2 ** Jfine * Bs
            masks = []
            for d in range(dim):
                nd = p.Bs[d] + 2 * g
                pos = (np.arange(nd) - g + int(grid.ixyz[b, d]) * p.Bs[d]) * s      # position in finest-grid units
                lim = gmin * s
                skip = (pos < lim) | (pos >= ntot - lim)
                idx = np.arange(nd)
                outside = (idx < g - gs) | (idx >= p.Bs[d] + g + gs)
                masks.append((skip, outside))
            shape = u[b, 0].shape
            sk = np.zeros(shape, bool)
            out = np.zeros(shape, bool)
            for d in range(dim):
                sh = [1, 1, 1]
                sh[2 - d] = -1
                sk |= masks[d][0].reshape(sh)
                out |= masks[d][1].reshape(sh)
            for c in range(2):
                got, exp = u[b, c], expected[b, c]
                bad += int(((got != -1.0) & out & ~sk).sum())                 # touched outside g_sync
                bad += int(((got != exp) & ~out & ~sk).sum())                 # wrong / missing inside g_sync
        assert bad == 0, (gs, bad)


@pytest.mark.parametrize("dim", [2, 3])
def test_host_forest_matches_independent_neighbor_search(dim):
    grid = two_level_grid(dim, refine=((0, 0, 0), (1, 0, 0)))
    f = Forest.from_blocks(dim, 2, grid.level, grid.ixyz, block_dist="sfc_z")
    hvy, lvl, ixyz, _ = f.active(0)
    g2 = O.Grid(level=lvl.astype(np.int64), ixyz=ixyz.astype(np.int64), dim=dim)
    ref = O.neighbor_table168(g2, 2)
    got = f.neighbors(0)[:, :g2.n]
    assert np.array_equal(got, ref)


def test_inverse_relation_is_an_involution_on_slots():
    L = O.lib()
    O.sync_ghosts_leaf   # ensure argtypes are set lazily below
    import ctypes as C
    L.orc_inverse_relation.argtypes = [C.c_int]
    L.orc_inverse_relation.restype = C.c_int
    for r in range(1, 169):
        inv = L.orc_inverse_relation(r)
        assert 1 <= inv <= 168 and L.orc_inverse_relation(inv) == r
        assert (inv - 1) // 56 == {0: 0, 1: 2, 2: 1}[(r - 1) // 56]
