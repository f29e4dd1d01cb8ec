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
@pytest.mark.parametrize("wavelet,Bs,ignore_filter", [("CDF40", 16, True), ("CDF44", 16, True), ("CDF20", 12, True), ("CDF62", 20, True),
                                                      ("CDF44", 16, False), ("CDF42", 16, False), ("CDF22", 12, False), ("CDF62", 20, False)])
def test_linear_field_sync_protocol(dim, wavelet, Bs, ignore_filter):
    """ignore_filter=True: sync_ghosts_RHS_tree (unit_test_Sync.f90:141); False: sync_ghosts_tree, whose restriction applies the HD
    filter of a lifted wavelet (unit_test_Sync.f90:136) -- the filter reproduces linear fields, so the protocol holds for both."""
    w = O.setup_wavelet(wavelet)
    g = w.g_default
    order = w.X
    g_rhs = max(w.X // 2, 1)
    p = O.Params(dim=dim, Bs=(Bs, Bs, Bs if dim == 3 else 1), g=g, g_rhs=g_rhs, n_eqn=2, Jmax=2)
    grid = two_level_grid(dim, refine=((0, 0, 0), (1, 1, 0)))
    nbr = O.neighbor_table168(grid, 2)
    interior = (slice(None), slice(None)) + O.interior(p)
    Jfine = 2
    # with the filter only the full depth is a critical case of the reference's test (unit_test_Sync.f90:205-210): at smaller depths
    # the filter reads same-level ghost nodes that stage 1 did not fill
    for gs in (range(g, g_rhs - 1, -1) if ignore_filter else (g,)):
        expected = O.alloc(grid, p)
        linear_field(grid, p, expected)
        u = np.full_like(expected, -1.0)
        u[interior] = expected[interior]
        n = O.sync_ghosts_leaf(grid, p, u, nbr, gs, gs, order, bool(w.lifted), ignore_filter=ignore_filter, w=w)
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


@pytest.mark.parametrize("wavelet,Bs,Jmax,seed", [("CDF40", 16, 3, 5), ("CDF44", 16, 3, 11), ("CDF62", 20, 3, 5)])
def test_sync_equals_geometric_definition_on_graded_grids(wavelet, Bs, Jmax, seed):
    """What the GPU path relies on (wabbit_b200/csrc/resolve.cuh): on a graded leaf grid, the reference's staged, table-driven
    synchronisation (ignore_Filter) gives every ghost point -- faces, edges and corners, full depth g -- the value of the
    coincident interior point of the leaf that owns it (same level or one finer), else the tensor-product interpolation
    (x, then y, then z) of the next-coarser lattice, whose points are again owned by leaves of that level or one finer."""
    from util import graded_blocks
    w = O.setup_wavelet(wavelet)
    g, order = w.g_default, w.X
    A = order // 2 - 1
    p = O.Params(dim=3, Bs=(Bs,) * 3, g=g, g_rhs=g, n_eqn=1, Jmax=Jmax)
    lv, ix = graded_blocks(3, 1, Jmax, seed, 0.3)
    grid = O.Grid(level=lv.astype(np.int64), ixyz=ix.astype(np.int64), dim=3)
    nbr = O.neighbor_table168(grid, Jmax)
    rng = np.random.default_rng(seed)
    u = O.alloc(grid, p, 1)
    u[:] = rng.standard_normal(u.shape)
    ref = u.copy()
    O.sync_ghosts_leaf(grid, p, ref, nbr, g, g, order, bool(w.lifted))
    look = grid.lookup()
    cf = {2: [.5, .5], 4: [-1 / 16, 9 / 16, 9 / 16, -1 / 16], 6: [3 / 256, -25 / 256, 150 / 256, 150 / 256, -25 / 256, 3 / 256]}[order]

    def lattice(L, P):
        n = 2 ** L * Bs
        P = [q % n for q in P]
        b = look.get((L,) + tuple(q // Bs for q in P))
        if b is not None:
            return u[b, 0, P[2] % Bs + g, P[1] % Bs + g, P[0] % Bs + g]
        b = look.get((L + 1,) + tuple((2 * q) // Bs for q in P))
        if b is None:
            return None
        return u[b, 0, (2 * P[2]) % Bs + g, (2 * P[1]) % Bs + g, (2 * P[0]) % Bs + g]

    def predicted(L, G, axis=2):
        if axis < 0:
            return lattice(L - 1, G)
        q = G[axis]
        c = list(G)
        if q % 2 == 0:
            c[axis] = q // 2
            return predicted(L, c, axis - 1)
        s = None
        for t in range(order):
            c[axis] = (q - 1) // 2 - A + t
            v = predicted(L, c, axis - 1)
            if v is None:
                return None
            s = cf[t] * v if s is None else s + cf[t] * v
        return s

    checked = {"copy": 0, "pred": 0}
    for b in range(grid.n):
        L = int(grid.level[b])
        for _ in range(40):
            i = rng.integers(-g, Bs + g, size=3)
            if all(0 <= q < Bs for q in i):
                continue
            G = [int(grid.ixyz[b, a]) * Bs + int(i[a]) for a in range(3)]
            v, kind = lattice(L, G), "copy"
            if v is None:
                v, kind = predicted(L, G), "pred"
            assert v is not None
            assert v == ref[b, 0, i[2] + g, i[1] + g, i[0] + g], (b, i, kind)
            checked[kind] += 1
    assert checked["copy"] > 100 and checked["pred"] > 100


def relation_directions(dim):
    """slot (1..56) -> direction (dx, dy, dz) of the neighbour table (neighborhood.f90:10-22)"""
    out = {}
    for dz in ((-1, 0, 1) if dim == 3 else (0,)):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                d = (dx, dy, dz)
                if d == (0, 0, 0):
                    continue
                nfree = 2 ** (sum(1 for v in d if v == 0) - (1 if dim == 2 else 0))
                for k in range(nfree):
                    out[O.same_level_code(d) + k] = d
    return out


@pytest.mark.parametrize("wavelet,Bs,Jmax,seed", [("CDF44", 16, 3, 11), ("CDF42", 16, 3, 3), ("CDF22", 16, 3, 7), ("CDF62", 20, 3, 5)])
def test_filtered_sync_equals_geometric_definition_on_graded_grids(wavelet, Bs, Jmax, seed):
    """What the GPU path relies on for sync_ghosts_tree with a lifted wavelet (wabbit_b200/csrc/fill.cuh, restrict_filter_kernel):
    a ghost point owned by a finer leaf F receives R(F) at the coincident point, where R(F) is the scaling coefficient of the
    one-level decomposition of F (HD in x, y, z with F's same-level neighbours in its ghosts) -- except inside the Nscl / Nscr strips
    that face F's coarser or finer neighbours, where it is the plain value; ghost points owned by a same-level leaf are copies and
    points owned by a coarser leaf are interpolated from the UNFILTERED coarse lattice (the coarse ghost nodes the stencil reaches
    lie in copy strips).  Compared bit for bit with the oracle's staged, table-driven restatement of the reference."""
    from util import graded_blocks
    w = O.setup_wavelet(wavelet)
    g, order = w.g_default, w.X
    A = order // 2 - 1
    p = O.Params(dim=3, Bs=(Bs,) * 3, g=g, g_rhs=g, n_eqn=1, Jmax=Jmax)
    lv, ix = graded_blocks(3, 1, Jmax, seed, 0.3)
    grid = O.Grid(level=lv.astype(np.int64), ixyz=ix.astype(np.int64), dim=3)
    nbr = O.neighbor_table168(grid, Jmax)
    rng = np.random.default_rng(seed)
    u = O.alloc(grid, p, 1)
    u[:] = rng.standard_normal(u.shape)
    ref = u.copy()
    O.sync_ghosts_leaf(grid, p, ref, nbr, g, g, order, True, ignore_filter=False, w=w)
    plain = u.copy()
    O.sync_ghosts_leaf(grid, p, plain, nbr, g, g, order, True)              # any fill of the non-same-level ghosts will do
    wd = np.zeros_like(u)
    O.fwt_tree(w, p, plain, wd)
    dirs = relation_directions(3)
    strips = []                                                            # per block: list of (lo[3], hi[3]) copy boxes, interior offsets
    for b in range(grid.n):
        boxes = set()
        for r in range(57, 169):
            if nbr[r - 1, b] >= 1:
                d = dirs[(r - 1) % 56 + 1]
                lo = tuple(0 if d[a] <= 0 else Bs - w.Nscr for a in range(3))
                hi = tuple(w.Nscl - 1 if d[a] < 0 else Bs - 1 for a in range(3))
                boxes.add((lo, hi))
        strips.append(sorted(boxes))
    look = grid.lookup()
    cf = {2: [.5, .5], 4: [-1 / 16, 9 / 16, 9 / 16, -1 / 16], 6: [3 / 256, -25 / 256, 150 / 256, 150 / 256, -25 / 256, 3 / 256]}[order]
    used = {"filtered": 0, "strip": 0}

    def lattice(L, P, filtered):
        n = 2 ** L * Bs
        P = [q % n for q in P]
        b = look.get((L,) + tuple(q // Bs for q in P))
        if b is not None:
            return u[b, 0, P[2] % Bs + g, P[1] % Bs + g, P[0] % Bs + g]
        b = look.get((L + 1,) + tuple((2 * q) // Bs for q in P))
        if b is None:
            return None
        o = [(2 * q) % Bs for q in P]
        if filtered and not any(all(lo[a] <= o[a] <= hi[a] for a in range(3)) for lo, hi in strips[b]):
            used["filtered"] += 1
            return wd[b, 0, o[2] + g, o[1] + g, o[0] + g]
        used["strip"] += filtered
        return u[b, 0, o[2] + g, o[1] + g, o[0] + g]

    def predicted(L, G, axis=2):
        if axis < 0:
            return lattice(L - 1, G, False)
        q = G[axis]
        c = list(G)
        if q % 2 == 0:
            c[axis] = q // 2
            return predicted(L, c, axis - 1)
        s = None
        for t in range(order):
            c[axis] = (q - 1) // 2 - A + t
            v = predicted(L, c, axis - 1)
            if v is None:
                return None
            s = cf[t] * v if s is None else s + cf[t] * v
        return s

    checked = {"copy": 0, "pred": 0}
    for b in range(grid.n):
        L = int(grid.level[b])
        for _ in range(60):
            i = rng.integers(-g, Bs + g, size=3)
            if all(0 <= q < Bs for q in i):
                continue
            G = [int(grid.ixyz[b, a]) * Bs + int(i[a]) for a in range(3)]
            v, kind = lattice(L, G, True), "copy"
            if v is None:
                v, kind = predicted(L, G), "pred"
            assert v is not None
            assert v == ref[b, 0, i[2] + g, i[1] + g, i[0] + g], (b, i, kind)
            checked[kind] += 1
    assert checked["copy"] > 100 and checked["pred"] > 100 and used["filtered"] > 20 and used["strip"] > 10
