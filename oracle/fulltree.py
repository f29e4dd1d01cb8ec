"""oracle/fulltree.py -- CPU restatement of WABBIT's adapt_tree with the full wavelet transformation on a graded leaf grid
(TEST INFRASTRUCTURE ONLY: imported by tests/, never by the product).

Reference (paths relative to the reference checkout):
  adapt_tree                                   LIB/MESH/adapt_tree.f90:11-260
  wavelet_decompose_full_tree                  LIB/MESH/adapt_tree.f90:268-545      (leaf-first and level-wise variants)
  wavelet_reconstruct_full_tree_CEoptimized    LIB/MESH/adapt_tree.f90:686-987
  init_full_tree / prune_fulltree2leafs        LIB/MESH/adapt_tree.f90:990-
  coarse_extension_modify                      LIB/MPI/reconstruction_step.f90:3-100
  sync_TMP_from_all / sync_TMP_from_MF / sync_SCWC_from_MC, prepare_ghost_synch_metadata
                                               LIB/MPI/synchronize_ghosts_generic.f90:1-175, 352-694
  sync_D2M                                     LIB/MESH/executeCoarsening_tree.f90:125-230
  coarseningIndicator_tree                     LIB/MESH/coarseningIndicator_tree.f90
  respectJmaxJmin_tree, ensureGradedness_tree, ensure_completeness_block   LIB/MESH/*.f90

Scope: lifted wavelets (useCoarseExtension = 1), useSecurityZone = 0 or 1 (addSecurityZone_CE_tree), indicator "threshold-state-vector" or "everywhere", periodic
domains, Bs >= Nrecon (no reconstruction of neighbours).  The tree is a dict keyed by (level, ix, iy, iz); every block carries the two
ghosted arrays of the reference, hvy_block (`blk`) and hvy_tmp (`tmp`), [nc, nz, ny, nx].

Facts of the reference's algorithm on a GRADED leaf grid that this restatement uses (each follows from the send / receive rules of
prepare_ghost_synch_metadata and is spelled out where it is used):
  (F1) a block that has daughters never has a coarser neighbour, and all its 3^d - 1 same-level neighbours exist in the full tree;
  (F2) with the coarse extension, a block's decomposition never depends on ghost nodes that face a coarser neighbour: the scaling
       coefficients within Nsc and the wavelet coefficients within Nwc of such a face are overwritten (copy / zero);
  (F3) in sync_SCWC_from_MC the ghost nodes a leaf receives from a coarser leaf are kept only at scaling positions, where the predictor
       returns the coincident coarse value; these coincident points are interior points of the sender.
"""
from __future__ import annotations

import ctypes as C
import functools
from typing import Dict, Optional, Tuple

import numpy as np

import oracle as O

Key = Tuple[int, int, int, int]
REF_STAY = 9      # REF_UNSIGNIFICANT_STAY, module_globals.f90:25-32


@functools.lru_cache(maxsize=None)
def dirs(dim):
    return [(dx, dy, dz) for dz in ((-1, 0, 1) if dim == 3 else (0,)) for dy in (-1, 0, 1) for dx in (-1, 0, 1) if (dx, dy, dz) != (0, 0, 0)]


def parent(k: Key) -> Key:
    return (k[0] - 1, k[1] >> 1, k[2] >> 1, k[3] >> 1)


def children(k: Key, dim: int):
    return [(k[0] + 1, 2 * k[1] + (c & 1), 2 * k[2] + ((c >> 1) & 1), 2 * k[3] + ((c >> 2) & 1) if dim == 3 else 0) for c in range(2 ** dim)]


def nbr_key(k: Key, d, dim: int) -> Key:
    m = (1 << k[0]) - 1                                   # periodic: modulo 2^level
    return (k[0], (k[1] + d[0]) & m, (k[2] + d[1]) & m, ((k[3] + d[2]) & m) if dim == 3 else 0)


class Tree:
    """Full tree: leaves plus all their ancestors down to Jmin (init_full_tree)."""

    def __init__(self, p: O.Params, w: O.Wavelet, grid: O.Grid, u: np.ndarray, Jmin: int, fd_half_width: int = 0):
        self.p, self.w, self.dim, self.Jmin = p, w, grid.dim, Jmin
        self.blk: Dict[Key, np.ndarray] = {}
        self.tmp: Dict[Key, np.ndarray] = {}
        self.leaf = set()
        for b in range(grid.n):
            k = (int(grid.level[b]),) + tuple(int(v) for v in grid.ixyz[b])
            self.blk[k] = u[b].copy()
            self.tmp[k] = np.zeros_like(u[b])
            self.leaf.add(k)
        for k in list(self.leaf):
            while k[0] > Jmin:
                k = parent(k)
                if k in self.blk:
                    break
                self.blk[k] = np.zeros_like(u[0])
                self.tmp[k] = np.zeros_like(u[0])
        self.decomposed = set()
        self.Nwcl = max(w.Nwcl, 2 * fd_half_width)        # setup_wavelet: widened to 2*FD_max_size (module_wavelets.f90:1404-1417)
        self.Nwcr = max(w.Nwcr, 2 * fd_half_width)
        self.Jmax_active = max(k[0] for k in self.leaf)

    # ------------------------------------------------------------------ geometry
    def interior(self):
        return (slice(None),) + O.interior(self.p)

    def is_leaf(self, k: Key) -> bool:
        return not any(c in self.blk for c in children(k, self.dim))

    def coarse_dirs(self, k: Key):
        """directions without a same-level block in the tree = the relations whose neighbour is coarser and that have no valid same-level or
        finer neighbour (reconstruction_step.f90:77-79); edge / corner regions owned by a coarser FACE neighbour carry no relation of their
        own in the reference, but their modify-patches are subsets of that face's patch, so listing them changes nothing"""
        return [d for d in dirs(self.dim) if nbr_key(k, d, self.dim) not in self.blk]

    # ------------------------------------------------------------------ ghost nodes from same-level blocks
    def sync_same_level(self, k: Key, src, gs: int):
        """stage 1 of sync_ghosts_generic for receiver k: copy the gs-deep strips of the same-level neighbours; src(nk) returns the array to
        read of neighbour nk or None (no such block / REF_TMP_EMPTY)"""
        p, dim, g = self.p, self.dim, self.p.g
        dst = self.blk[k]
        for d in dirs(dim):
            s = src(nbr_key(k, d, dim))
            if s is None:
                continue
            rs, ss = [slice(None)], [slice(None)]
            for a in (2, 1, 0):                     # array axes z, y, x
                if a >= dim:
                    rs.append(slice(0, 1))
                    ss.append(slice(0, 1))
                    continue
                B = p.Bs[a]
                if d[a] < 0:
                    rs.append(slice(g - gs, g))
                    ss.append(slice(g + B - gs, g + B))
                elif d[a] > 0:
                    rs.append(slice(g + B, g + B + gs))
                    ss.append(slice(g, g + gs))
                else:
                    rs.append(slice(g, g + B))
                    ss.append(slice(g, g + B))
            dst[tuple(rs)] = s[tuple(ss)]

    # ------------------------------------------------------------------ block operations
    def fwt(self, k: Key):
        """hvy_tmp = hvy_block (with synchronised ghosts); hvy_block = decomposition (adapt_tree.f90:424-446)"""
        L = O._wl()
        self.tmp[k] = self.blk[k].copy()
        out = self.blk[k].copy()      # waveletDecomposition_optimized_block writes the interior only: ghosts keep their values
        L.orc_fwt_block(C.byref(self.w), self.dim, self.p.g, O._bs(self.p.Bs), out.shape[0], O._p(self.tmp[k]), O._p(out))
        self.blk[k] = out
        self.decomposed.add(k)

    def ce_modify(self, k: Key, clear_wc=True, copy_sc=True):
        L = O.lib()
        O.coarse_extension_modify  # noqa: B018  (argtypes of orc_ce_modify_block are set lazily there)
        if not hasattr(L, "_ce_ready"):
            L.orc_ce_modify_block.argtypes = [C.c_int, C.c_int, O._ip, C.c_int, O._dp, O._dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
            L._ce_ready = True
        n = 0
        for d in self.coarse_dirs(k):
            rel = O.same_level_code(d) + 56
            L.orc_ce_modify_block(self.dim, self.p.g, O._bs(self.p.Bs), self.blk[k].shape[0], O._p(self.blk[k]), O._p(self.tmp[k]), rel,
                                  self.Nwcl, self.Nwcr, self.w.Nscl, self.w.Nscr, int(clear_wc), int(copy_sc))
            n += 1
        return n

    def d2m(self, level: int):
        """sync_D2M(sync_case="level"): the scaling coefficients of every block on `level` become an octant of its mother's hvy_block"""
        p, g, dim = self.p, self.p.g, self.dim
        for k in [k for k in self.blk if k[0] == level and k[0] > self.Jmin]:
            m = parent(k)
            if m not in self.blk:
                continue
            q = (k[1] & 1, k[2] & 1, k[3] & 1)
            h = [p.Bs[a] // 2 for a in range(3)]
            sx = slice(g + q[0] * h[0], g + (q[0] + 1) * h[0])
            sy = slice(g + q[1] * h[1], g + (q[1] + 1) * h[1])
            if dim == 3:
                sz = slice(g + q[2] * h[2], g + (q[2] + 1) * h[2])
                self.blk[m][:, sz, sy, sx] = self.blk[k][:, g:g + p.Bs[2]:2, g:g + p.Bs[1]:2, g:g + p.Bs[0]:2]
            else:
                self.blk[m][:, :, sy, sx] = self.blk[k][:, :, g:g + p.Bs[1]:2, g:g + p.Bs[0]:2]


def decompose_full_tree(p: O.Params, w: O.Wavelet, grid: O.Grid, u: np.ndarray, Jmin: int = 1, fd_half_width: int = 0,
                        force_leaf_first: Optional[bool] = None, use_coarse_extension: Optional[bool] = None) -> Tree:
    """wavelet_decompose_full_tree (adapt_tree.f90:268-545).  leaf-first (all Bs >= 3*max|HD tap index|): every leaf is decomposed in the
    first pass after a full synchronisation (sync_TMP_from_all: same level, restriction through the HD filter, prediction), then the
    mothers level by level; level-wise otherwise: per level, all blocks of the level together after sync_TMP_from_MF."""
    t = Tree(p, w, grid, u, Jmin, fd_half_width)
    dim = grid.dim
    # params%useCoarseExtension: default isLiftedWavelet (ini_file_to_params.f90:543); an .ini may switch it on for an unlifted wavelet
    # (TESTING/acm/3vortices/3vorticesAdaptFD4_CDF40), which then takes the very same code path with Nsc = 0
    use_ce = bool(w.lifted) if use_coarse_extension is None else bool(use_coarse_extension)
    t.use_ce = use_ce
    F = max(abs(w.hd_lo), w.hd_hi)
    leaf_first = all(p.Bs[a] >= 3 * F for a in range(dim)) if force_leaf_first is None else force_leaf_first
    t.leaf_first = leaf_first
    gs = max(abs(w.hd_lo), w.gd_hi)                         # g_this, adapt_tree.f90:403
    if leaf_first:
        # iteration 0: all leaves carry -1, mothers REF_TMP_EMPTY (they neither send nor receive): a leaf-grid synchronisation
        nbr = O.neighbor_table168(grid, max(int(grid.level.max()), 1) + 8)
        hv = u.copy()
        O.sync_ghosts_leaf(grid, p, hv, nbr, gs, gs, w.X, bool(w.lifted), ignore_filter=not w.lifted, w=w)
        for b in range(grid.n):
            k = (int(grid.level[b]),) + tuple(int(v) for v in grid.ixyz[b])
            t.blk[k] = hv[b].copy()
        for k in sorted(t.leaf):
            t.fwt(k)
        for k in sorted(t.leaf):
            if use_ce:                                      # adapt_tree.f90:479
                t.ce_modify(k)                              # CE_case="ref", s_ref=-1: leaves only (reconstruction_step.f90:66)
    level = t.Jmax_active
    while level >= Jmin:
        if leaf_first:
            todo = [k for k in t.blk if k[0] == level and k not in t.leaf]       # mothers of this level (REF_TMP_EMPTY -> -1)
        else:
            todo = [k for k in t.blk if k[0] == level]                           # leaves and mothers of this level
        if todo:
            # sync_TMP_from_MF: same-level senders -- hvy_tmp if already decomposed (status 0), hvy_block if in this pass (-1); finer
            # blocks have a non-empty mother on this level by now and do not send (prepare_ghost_synch_metadata: non-root blocks do not
            # send to coarser neighbours); coarser neighbours are not asked (F2)
            todo_set = set(todo)

            def src(nk):
                if nk not in t.blk:
                    return None
                if nk in todo_set:
                    return snapshot[nk]
                return t.tmp[nk] if nk in t.decomposed else None
            snapshot = {k: t.blk[k].copy() for k in todo}                        # senders are read before anyone is decomposed
            for k in sorted(todo):
                t.sync_same_level(k, src, gs)
            for k in sorted(todo):
                t.fwt(k)
            for k in sorted(todo):
                if k in t.leaf and use_ce:
                    t.ce_modify(k)
        t.d2m(level)
        level -= 1
    return t


def threshold_full_tree(t: Tree, eps: float, norm=None, eps_norm: str = "Linfty", thresh_comp=None, level_ref: int = 0,
                        force_maxlevel_dealiasing: bool = False, indicator: str = "threshold-state-vector") -> Dict[Key, int]:
    """coarseningIndicator_tree on the decomposed full tree: every block (leaf or not) gets -1 / 0 from its own coefficients."""
    L = O._wl()
    p = t.p
    st = {}
    if indicator == "everywhere":
        return {k: (-1 if k in t.leaf else 0) for k in t.blk}
    nc = next(iter(t.blk.values())).shape[0]
    tc = np.ascontiguousarray(np.ones(nc) if thresh_comp is None else thresh_comp, dtype=np.int32)
    e = np.full(nc, eps, dtype=np.float64)
    nrm = None if norm is None else np.ascontiguousarray(norm, dtype=np.float64)
    det = np.zeros(nc)
    for k in t.blk:
        if force_maxlevel_dealiasing and k[0] == level_ref:
            st[k] = -1
            continue
        st[k] = int(L.orc_threshold_block(p.dim, p.g, O._bs(p.Bs), nc, O._p(t.blk[k]), k[0], level_ref, O.EPS_NORMS[eps_norm],
                                          tc.ctypes.data_as(O._ip), O._p(e), O._p(nrm), O._p(det)))
    return st


# ----------------------------------------------------------------------------------------------------------------------
# grid decision: respectJmaxJmin_tree + ensureGradedness_tree(check_daughters=.true.) on the full tree
# ----------------------------------------------------------------------------------------------------------------------
def finer_neighbors(t: Tree, k: Key):
    """blocks one level finer that touch block k (relations 113..168 of the full tree's hvy_neighbor)"""
    out = []
    for d in dirs(t.dim):
        nk = nbr_key(k, d, t.dim)
        for c in children(nk, t.dim):
            ok = True
            for a in range(t.dim):
                off = c[1 + a] & 1
                if (d[a] > 0 and off != 0) or (d[a] < 0 and off != 1):
                    ok = False
            if ok and c in t.blk:
                out.append(c)
    return out


def security_zone(t: Tree, st: Dict[Key, int], eps: float, norm=None, eps_norm: str = "Linfty", thresh_comp=None, level_ref: int = 0,
                  force_maxlevel_dealiasing: bool = False) -> Dict[Key, int]:
    """addSecurityZone_CE_tree (LIB/MESH/securityZone_tree.f90:140-298): for every significant block (status 0) and every same-level
    neighbour with -1, threshold the significant block's coefficients inside the Nwcl / Nwcr deep strip facing that neighbour
    (get_indices_of_modify_patch); if they are significant the neighbour is kept (status 0)."""
    L = O.lib()
    if not hasattr(L, "_tbb_ready"):
        L.orc_threshold_block_box.argtypes = [C.c_int, C.c_int, O._ip, C.c_int, O._dp, C.c_int, C.c_int, C.c_int, O._ip, O._dp, O._dp, O._dp, O._ip, O._ip]
        L.orc_threshold_block_box.restype = C.c_int
        L._tbb_ready = True
    p, dim = t.p, t.dim
    out = dict(st)
    nc = next(iter(t.blk.values())).shape[0]
    tc = np.ascontiguousarray(np.ones(nc) if thresh_comp is None else thresh_comp, dtype=np.int32)
    e = np.full(nc, eps, dtype=np.float64)
    nrm = None if norm is None else np.ascontiguousarray(norm, dtype=np.float64)
    det = np.zeros(nc)
    for k in t.blk:
        if st[k] != 0 or (force_maxlevel_dealiasing and k[0] == level_ref):
            continue
        for d in dirs(dim):
            nk = nbr_key(k, d, dim)
            if nk not in t.blk or st[nk] != -1:
                continue
            lo = np.zeros(3, np.int32)
            hi = np.zeros(3, np.int32)
            for a in range(dim):
                B = p.Bs[a]
                lo[a] = max(B - t.Nwcr, 0) if d[a] > 0 else 0
                hi[a] = min(t.Nwcl, B) - 1 if d[a] < 0 else B - 1
            r = L.orc_threshold_block_box(dim, p.g, O._bs(p.Bs), nc, O._p(t.blk[k]), k[0], level_ref, O.EPS_NORMS[eps_norm], tc.ctypes.data_as(O._ip),
                                          O._p(e), O._p(nrm), O._p(det), lo.ctypes.data_as(O._ip), hi.ctypes.data_as(O._ip))
            if r == 0:
                out[nk] = 0
    return out


def decide(t: Tree, st: Dict[Key, int], Jmin: int) -> Dict[Key, int]:
    """Which blocks are deleted (-1).  respectJmaxJmin_tree (blocks on Jmin stay); then, until nothing changes
    (ensureGradedness_tree.f90, ensure_completeness_block.f90, statuses only ever move from -1 to "stay", so the result does not depend
    on the order of the sweep): a block keeps -1 only if all its 2^d sisters have -1 (completeness), none of its daughters stays
    (check_daughters), and no finer neighbour stays (gradedness)."""
    st = dict(st)
    for k in st:
        if st[k] == -1 and k[0] <= Jmin:
            st[k] = REF_STAY
    changed = True
    while changed:
        changed = False
        for k in sorted(st):
            if st[k] != -1:
                continue
            sis = children(parent(k), t.dim)
            stay = any(s not in st or st[s] != -1 for s in sis)
            if not stay and not t.is_leaf(k):
                stay = any(c in st and st[c] != -1 for c in children(k, t.dim))
            if not stay:
                stay = any(st[f] != -1 for f in finer_neighbors(t, k))
            if stay:
                st[k] = REF_STAY
                changed = True
    return st


# ----------------------------------------------------------------------------------------------------------------------
# adapt_tree: decomposition, indicator, deletion, coarse extension on the new interfaces, reconstruction
# ----------------------------------------------------------------------------------------------------------------------
def ndep2(w: O.Wavelet, fd_half_width: int):
    """Nrecon and Ndep2 of setup_wavelet incl. the FD widening (module_wavelets.f90:1368-1417)"""
    dl = max(2 * fd_half_width - w.Nwcl, 0)
    dr = max(2 * fd_half_width - w.Nwcr, 0)
    nrl, nrr = w.Nreconl + dl, w.Nreconr + dr
    d2l = w.Nreconl + max((abs(w.hr_lo) + 1) // 2 - 1, 0) + dl
    d2r = w.Nreconr + (w.hr_hi + 1) // 2 + dr
    return nrl, nrr, d2l, d2r


def _fill_from_coarse(t: Tree, k: Key, d):
    """ghost patch of direction d of block k from its coarser leaf neighbour in sync_SCWC_from_MC + coarse_extension_modify (F3): scaling
    positions take the coincident value of the sender's hvy_tmp, everything else is a wavelet coefficient and is set to zero"""
    p, dim, g = t.p, t.dim, t.p.g
    ck = parent(nbr_key(k, d, dim))
    src = t.tmp[ck]
    dst = t.blk[k]
    rng_f, rng_c = [], []
    for a in range(3):
        if a >= dim:
            rng_f.append(np.array([0]))
            continue
        B = p.Bs[a]
        loc = np.arange(-g, 0) if d[a] < 0 else (np.arange(B, B + g) if d[a] > 0 else np.arange(0, B))
        rng_f.append(loc)
    n = 2 ** k[0]
    idx = []
    for a in range(dim):
        B = p.Bs[a]
        glob = (k[1 + a] * B + rng_f[a]) % (n * B)                    # global fine coordinate, periodic
        even = glob % 2 == 0
        cloc = glob // 2 - ck[1 + a] * B                               # coordinate inside the coarse sender
        idx.append((rng_f[a] + g, even, cloc + g))
    if dim == 3:
        (fx, ex, cx), (fy, ey, cy), (fz, ez, cz) = idx
        dst[:, fz[:, None, None], fy[None, :, None], fx[None, None, :]] = 0.0
        assert ((cx[ex] >= g) & (cx[ex] < g + p.Bs[0])).all() and ((cy[ey] >= g) & (cy[ey] < g + p.Bs[1])).all() and \
            ((cz[ez] >= g) & (cz[ez] < g + p.Bs[2])).all()          # F3: interior points of the sender
        dst[:, fz[ez][:, None, None], fy[ey][None, :, None], fx[ex][None, None, :]] = \
            src[:, cz[ez][:, None, None], cy[ey][None, :, None], cx[ex][None, None, :]]
    else:
        (fx, ex, cx), (fy, ey, cy) = idx
        dst[:, 0, fy[:, None], fx[None, :]] = 0.0
        assert ((cx[ex] >= g) & (cx[ex] < g + p.Bs[0])).all() and ((cy[ey] >= g) & (cy[ey] < g + p.Bs[1])).all()
        dst[:, 0, fy[ey][:, None], fx[ex][None, :]] = src[:, 0, cy[ey][:, None], cx[ex][None, :]]


def adapt_tree(p: O.Params, w: O.Wavelet, grid: O.Grid, u: np.ndarray, eps: float, Jmin: int = 1, norm=None, eps_norm: str = "Linfty",
               thresh_comp=None, level_ref: int = 0, force_maxlevel_dealiasing: bool = False, indicator: str = "threshold-state-vector",
               fd_half_width: int = 0, force_leaf_first: Optional[bool] = None, use_security_zone: bool = False,
               use_coarse_extension: Optional[bool] = None, mask_keeps=None):
    """adapt_tree (adapt_tree.f90:11-260) for a lifted wavelet with the coarse extension, with or without the security zone.  Returns
    (new grid, new data [nb, nc, nz, ny, nx] with meaningful interiors, info dict)."""
    dim = grid.dim
    t = decompose_full_tree(p, w, grid, u, Jmin, fd_half_width, force_leaf_first, use_coarse_extension)
    st0 = threshold_full_tree(t, eps, norm, eps_norm, thresh_comp, level_ref, force_maxlevel_dealiasing, indicator)
    if mask_keeps is not None and indicator != "everywhere":
        # threshold_mask: coarseningIndicatorMask_tree (coarseningIndicator_tree.f90:290-331) -- a block whose mask function is not constant
        # over its interior stays (status max(status, 0)); blocks on Jmax under force_maxlevel_dealiasing are not asked
        for k in st0:
            if force_maxlevel_dealiasing and k[0] == level_ref:
                continue
            if st0[k] == -1 and mask_keeps(k):
                st0[k] = 0
    if use_security_zone and indicator != "everywhere":
        st0 = security_zone(t, st0, eps, norm, eps_norm, thresh_comp, level_ref, force_maxlevel_dealiasing)
    st = decide(t, st0, Jmin)
    deleted = {k for k in st if st[k] == -1}
    for k in deleted:                                                   # "any block with -1 can simply be deleted"
        del t.blk[k]
        del t.tmp[k]
    leaves = {k for k in t.blk if t.is_leaf(k)}
    marked = sorted(k for k in leaves if t.coarse_dirs(k))              # leaves at a coarse/fine interface of the NEW grid
    if not t.use_ce:                                                    # no coarse extension: "restore original values" (adapt_tree.f90:236-241)
        keys = sorted(leaves)
        new_grid = O.Grid(level=np.array([k[0] for k in keys], dtype=np.int64), ixyz=np.array([k[1:] for k in keys], dtype=np.int64), dim=dim)
        return new_grid, np.stack([t.tmp[k] for k in keys]), {"status0": st0, "status": st, "marked": [], "leaf_only": True,
                                                              "leaf_first": t.leaf_first, "n_deleted": len(deleted)}
    nrl, nrr, d2l, d2r = ndep2(w, fd_half_width)
    if any(p.Bs[a] < max(nrl, nrr) for a in range(dim)):
        # Bs < Nrecon (adapt_tree.f90:771-803, reconstruct_neighbors): the modified coefficients of an interface block reach into its same-level
        # neighbours' reconstruction, so the leaves that have a marked same-level neighbour are reconstructed as well
        mset = set(marked)
        extra = [k for k in sorted(leaves) if k not in mset and any(nbr_key(k, d, dim) in mset for d in dirs(dim))]
        marked = sorted(mset | set(extra))
    leaf_only = all(p.Bs[a] >= d2l and p.Bs[a] >= d2r for a in range(dim))
    # coarse extension on the lasting interfaces: wavelet coefficients only (adapt_tree.f90:222-228)
    for k in marked:
        t.ce_modify(k, clear_wc=True, copy_sc=False)
    levels = [None] if leaf_only else list(range(Jmin, max(k[0] for k in t.blk) + 1))
    for level in levels:
        todo = [k for k in marked if level is None or k[0] == level]
        # sync_SCWC_from_MC: coefficients of the same-level neighbours (leaves or mothers), coarse values at scaling positions
        for k in todo:
            t.sync_same_level(k, lambda nk: t.blk.get(nk), p.g)
            for d in t.coarse_dirs(k):
                _fill_from_coarse(t, k, d)
        for k in todo:
            t.ce_modify(k, clear_wc=True, copy_sc=False)
        L = O._wl()
        for k in todo:
            out = t.blk[k].copy()
            L.orc_iwt_block(C.byref(w), dim, p.g, O._bs(p.Bs), out.shape[0], O._p(t.blk[k]), O._p(out))
            t.tmp[k] = out                                            # waveletReconstruction_optimized_block(hvy_block -> hvy_tmp)
        for k in todo:
            t.blk[k] = t.tmp[k].copy()                                # hvy_block = hvy_tmp
        for k in t.blk:                                               # blocks of this level that are not reconstructed: original values
            if (level is None or k[0] == level) and k not in todo:
                t.blk[k] = t.tmp[k].copy()
    keys = sorted(leaves)                                             # prune_fulltree2leafs
    new_grid = O.Grid(level=np.array([k[0] for k in keys], dtype=np.int64), ixyz=np.array([k[1:] for k in keys], dtype=np.int64), dim=dim)
    data = np.stack([t.blk[k] for k in keys])
    return new_grid, data, {"status0": st0, "status": st, "marked": marked, "leaf_only": leaf_only, "leaf_first": t.leaf_first,
                            "n_deleted": len(deleted)}
