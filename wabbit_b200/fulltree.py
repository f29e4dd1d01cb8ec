"""adapt_tree with the full wavelet transformation on the device (single rank): lifted wavelets with the coarse extension, and unlifted
ones (no coarse extension: the decomposition only drives the indicator, values are kept).

The reference decomposes the whole tree -- leaves and all their ancestors ("mothers") -- from fine to coarse
(wavelet_decompose_full_tree, LIB/MESH/adapt_tree.f90:268-545), thresholds every block (coarseningIndicator_tree), decides on the
light data which blocks go, and reconstructs the leaves that end up next to a coarser neighbour from their coefficients with the
coarse extension applied (wavelet_reconstruct_full_tree_CEoptimized, :686-987).  Here the mothers become extra RESIDENT blocks in free
slots of the device arrays, and every heavy-data step is one of the existing entry points of include/wabbit_gpu.h run on a block list
with its own neighbour table:

   decomposition of a set of blocks with ghost nodes from same-level blocks   wgpu_set_treecodes + wgpu_set_topology + wgpu_fwt
   coarse extension on the leaves of that set                                  wgpu_coarse_extension
   scaling coefficients -> octants of the mothers (sync_D2M)                   wgpu_coarsen
   details and refinement flags of the set                                     wgpu_threshold

hvy_block of the reference <-> array W = hvy_work(:,:,:,:,:,2) (decomposed values); hvy_tmp <-> array U = hvy_block here (original values
of the leaves, assembled scaling coefficients of the mothers).  Two variants, as in the reference: leaf-first when Bs >= 3*max|HD tap|
(all leaves in one pass after the full synchronisation with the filtered restriction, then the mothers level by level) and level-wise
otherwise (per level, leaves and mothers of the level together; a leaf takes the ghost nodes that face finer blocks from their mother).

Light data (which blocks exist, their slots, the per-pass neighbour tables) are host logic: numpy + libwabbit_host.so.
"""
from __future__ import annotations

import ctypes as C
import os
import time
from typing import Dict, List, Optional, Tuple

import numpy as np

from ._native import host_lib
from .forest import Forest

Key = Tuple[int, int, int, int]
HVY_BLOCK, HVY_WORK = 0, 1
WD = (HVY_WORK, 2)


def _dirs(dim):
    return [(dx, dy, dz) for dz in ((-1, 0, 1) if dim == 3 else (0,)) for dy in (-1, 0, 1) for dx in (-1, 0, 1) if (dx, dy, dz) != (0, 0, 0)]


def _code(d) -> int:
    """same-level slot (1..56) of a direction (find_neighbor, LIB/MESH/find_neighbors.f90:60-95)"""
    nzero = sum(1 for v in d if v == 0)
    if nzero == 2:
        code = 1
        for i in range(3):
            if d[i] != 0:
                code += 8 * i
            if d[i] == 1:
                code += 4
        return code
    if nzero == 1:
        code, apply_free = 25, 1
        for i in range(3):
            if d[i] == 0:
                code += 8 * (2 - i)
            else:
                if d[i] == 1:
                    code += apply_free * 2
                apply_free += 1
        return code
    return 49 + sum(1 << i for i in range(3) if d[i] == 1)


def _parent(k: Key) -> Key:
    return (k[0] - 1, k[1] >> 1, k[2] >> 1, k[3] >> 1)


def _children(k: Key, dim: int) -> List[Key]:
    """digit order of the treecode: bit0 -> y, bit1 -> x, bit2 -> z (refinementExecute.f90)"""
    return [(k[0] + 1, 2 * k[1] + ((c >> 1) & 1), 2 * k[2] + (c & 1), 2 * k[3] + ((c >> 2) & 1) if dim == 3 else 0) for c in range(2 ** dim)]


def _pack(level, ixyz):
    """sortable int64 code of a block position (level, ix, iy, iz); arrays in, array out"""
    level = np.asarray(level, dtype=np.int64)
    ixyz = np.asarray(ixyz, dtype=np.int64).reshape(-1, 3)
    return (level << 57) | (ixyz[:, 2] << 38) | (ixyz[:, 1] << 19) | ixyz[:, 0]


def _build_tree(dim: int, Jmin: int, level: np.ndarray, pos: np.ndarray):
    """leaves + all ancestors down to Jmin, sorted by position code: (level[n], pos[n,3], leaf_of[n]) as int64 / index into the input or -1
    (libwabbit_host.so: whost_ft_build)"""
    lv = np.ascontiguousarray(level, dtype=np.int32)
    px = np.ascontiguousarray(pos, dtype=np.int32).reshape(-1, 3)
    cap = len(lv) + len(lv) // 4 + 64
    i32 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    while True:
        lo, po, lf = np.zeros(cap, np.int32), np.zeros((cap, 3), np.int32), np.zeros(cap, np.int32)
        n = C.c_int32()
        rc = host_lib().whost_ft_build(dim, Jmin, len(lv), i32(lv), i32(px), cap, C.byref(n), i32(lo), i32(po), i32(lf))
        if rc == 2:
            cap *= 2
            continue
        if rc:
            raise RuntimeError(f"whost_ft_build: {rc}")
        n = n.value
        return lo[:n].astype(np.int64), po[:n].astype(np.int64), lf[:n].astype(np.int64)


def _encode_treecodes(dim: int, level: np.ndarray, ixyz: np.ndarray, Jmax: int) -> np.ndarray:
    """numerical binary treecode (module_treelib.f90:837-871): digit bit0 <- y, bit1 <- x, bit2 <- z; bit i of a coordinate goes to
    digit i + Jmax - level (libwabbit_host.so: whost_encode_many)"""
    lv = np.ascontiguousarray(level, dtype=np.int32)
    px = np.ascontiguousarray(ixyz, dtype=np.int32).reshape(-1, 3)
    out = np.zeros(len(lv), np.int64)
    i32 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    if host_lib().whost_encode_many(dim, Jmax, len(lv), i32(lv), i32(px), out.ctypes.data_as(C.POINTER(C.c_int64))):
        raise RuntimeError("whost_encode_many failed")
    return out


def _encode_treecodes_numpy(dim: int, level: np.ndarray, ixyz: np.ndarray, Jmax: int) -> np.ndarray:
    """the same in numpy (reference formulation for tests/test_host.py)"""
    level = level.astype(np.int64)
    p = [ixyz[:, 1].astype(np.int64), ixyz[:, 0].astype(np.int64), ixyz[:, 2].astype(np.int64)]
    tc = np.zeros(len(level), dtype=np.int64)
    for i in range(Jmax):
        use = i < level
        sh = (i + Jmax - level) * dim
        for d in range(dim):
            tc |= np.where(use, ((p[d] >> i) & 1) << np.where(use, sh + d, 0), 0)
    return tc


class FullTree:
    """Leaves of `forest` (slots = their hvy ids) plus all ancestors down to Jmin in free slots behind them (init_full_tree).
    Light data are numpy arrays over the blocks of the tree, sorted by position code; `slot` / `leaf` give dict / set views."""

    def __init__(self, sol, forest: Forest, Jmin: int = 1):
        self.sol, self.forest, self.dim, self.Jmin = sol, forest, forest.dim, Jmin
        dim = self.dim
        hvy, lvl, ixyz, _ = forest.active(0)
        # init_full_tree: leaves (slots = their hvy ids) + ancestors; mothers get the free slots behind the leaves, in position order
        level, pos, leaf_of = _build_tree(dim, Jmin, lvl, ixyz)
        is_leaf = leaf_of >= 0
        slots = np.zeros(len(level), np.int64)
        slots[is_leaf] = hvy[leaf_of[is_leaf]]
        slots[~is_leaf] = int(hvy.max()) + 1 + np.arange(int((~is_leaf).sum()))
        if slots.max() > sol.max_blocks:
            raise MemoryError(f"full tree needs {int(slots.max())} block slots, max_blocks = {sol.max_blocks}")
        self._set_blocks(level, pos, slots, is_leaf, presorted=True)
        self.Jmax_active = int(lvl.max())
        self.st = np.zeros(len(level), np.int32)
        self.det = None
        self.timing = {} if os.environ.get("WABBIT_FT_TIMING") else None
        F = sol.wavelet_filter_width()
        p = sol.params
        self.leaf_first = all(p.Bs[a] >= 3 * F for a in range(dim))
        # useCoarseExtension: default isLiftedWavelet (ini_file_to_params.f90:543); an .ini may switch it on for an unlifted wavelet, which
        # then takes the same path with Nsc = 0 (TESTING/acm/3vortices/3vorticesAdaptFD4_CDF40)
        self.lifted = (p.wavelet[4] != "0") if p.useCoarseExtension < 0 else bool(p.useCoarseExtension)

    def _set_blocks(self, level, pos, slots, is_leaf, presorted: bool = False):
        code = _pack(level, pos)
        o = np.arange(len(code)) if presorted else np.argsort(code)
        self.code, self.level, self.pos, self.slots, self.is_leaf = code[o], level[o], pos[o], slots[o], is_leaf[o]
        self._build_tables()
        return o

    def _build_tables(self):
        """same-level neighbour index per direction, mother index and daughter indices of every block of the tree (or -1): the whole
        neighbourhood logic of the passes and of the grid decision reads these (libwabbit_host.so: whost_ft_tables)"""
        dim, n = self.dim, len(self.code)
        self.dirs = _dirs(dim)
        self._lvl32 = np.ascontiguousarray(self.level, dtype=np.int32)
        self._pos32 = np.ascontiguousarray(self.pos, dtype=np.int32)
        nb = np.zeros((n, len(self.dirs)), np.int32)
        par = np.zeros(n, np.int32)
        child = np.zeros((n, 2 ** dim), np.int32)
        i32 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        rc = host_lib().whost_ft_tables(dim, n, i32(self._lvl32), i32(self._pos32), i32(nb), i32(par), i32(child))
        if rc:
            raise RuntimeError(f"whost_ft_tables: {rc}")
        self.nb, self.par, self.child = nb, par, child
        self._rows_ready = False

    def _upload_rows(self):
        """hvy_neighbor rows of EVERY block of the tree, built once per tree state: same-level relations to whatever block sits there
        (leaf or mother); for leaves, directions without a same-level block become coarser relations (slot + 56), which is where the
        coarse extension acts.  All blocks are registered as data sources (wgpu_set_treecodes); a pass then only names its active list."""
        sol, dim = self.sol, self.dim
        ld = int(self.slots.max())
        # one (168, max_blocks) table per solver, reused by every tree: rows 112..167 (finer relations) are never set here and stay -1
        N = max(int(getattr(sol, "max_blocks", ld)), ld)
        nbr = getattr(sol, "_ft_rows", None)
        if nbr is None or nbr.shape[1] != N:
            nbr = np.full((168, N), -1, dtype=np.int32)
            try:
                sol._ft_rows = nbr
            except AttributeError:
                pass
        else:
            nbr[:112, :getattr(sol, "_ft_rows_used", N)] = -1
        try:
            sol._ft_rows_used = ld
        except AttributeError:
            pass
        leaf = np.ascontiguousarray(self.is_leaf & (self.level > 0), dtype=np.int32)
        codes = np.array([_code(d) for d in self.dirs], dtype=np.int32)
        slots32 = np.ascontiguousarray(self.slots, dtype=np.int32)
        i32 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        rc = host_lib().whost_ft_rows(dim, len(self.code), i32(self._lvl32), i32(self._pos32), i32(self.nb), i32(slots32), i32(leaf), i32(codes),
                                      nbr.shape[1], i32(nbr))
        if rc:
            raise RuntimeError(f"whost_ft_rows: {rc}")
        self._rows = nbr
        tc = _encode_treecodes(dim, self.level, self.pos, self.forest.Jmax)
        sol.set_treecodes(self.slots.astype(np.int32), self._lvl32, tc)
        self._rows_ready = True

    def _rows_numpy(self, ld: int) -> np.ndarray:
        """the table of _upload_rows in numpy (reference formulation of whost_ft_rows, used by tests/test_host.py)"""
        nbr = np.full((168, ld), -1, dtype=np.int32)
        col = self.slots - 1
        leaf = self.is_leaf & (self.level > 0)
        for q, d in enumerate(self.dirs):
            j = self.nb[:, q]
            hit = j >= 0
            nbr[_code(d) - 1, col[hit]] = self.slots[j[hit]]
            miss = np.flatnonzero(~hit & leaf)
            if len(miss):
                c = self._find(self.level[miss] - 1, self._neighbor_pos(miss, d) >> 1)
                ok = c >= 0
                nbr[_code(d) - 1 + 56, col[miss[ok]]] = self.slots[c[ok]]
        return nbr

    # ---- views used by callers and tests
    def keys(self, idx=None):
        idx = np.arange(len(self.code)) if idx is None else idx
        return [(int(l), int(x[0]), int(x[1]), int(x[2])) for l, x in zip(self.level[idx], self.pos[idx])]

    @property
    def slot(self) -> Dict[Key, int]:
        return dict(zip(self.keys(), (int(v) for v in self.slots)))

    @property
    def leaf(self):
        return set(self.keys(np.flatnonzero(self.is_leaf)))

    def _find(self, level, pos):
        """index into the tree arrays of the blocks at (level, pos) or -1"""
        q = _pack(level, pos)
        i = np.searchsorted(self.code, q)
        i = np.minimum(i, len(self.code) - 1)
        return np.where(self.code[i] == q, i, -1)

    def _neighbor_pos(self, idx, d):
        n = (1 << self.level[idx])[:, None]
        dd = np.array([d[a] if a < self.dim else 0 for a in range(3)], dtype=np.int64)[None, :]
        return (self.pos[idx] + dd) % n

    # ------------------------------------------------------------------ per-pass topology
    def _tick(self, name, t0):
        if self.timing is not None:
            self.sol.synchronize()
            self.timing[name] = self.timing.get(name, 0.0) + (time.perf_counter() - t0)
        return time.perf_counter()

    def set_pass_topology(self, idx: np.ndarray):
        """a pass = a list of active blocks (indices into the tree arrays).  The whole tree -- leaves and mothers -- is registered on the
        device once per tree state (wgpu_set_grid: position hash, neighbour relations derived on the GPU); a pass then only names its
        active list (wgpu_set_active).  Returns the indices in the order of the active list (ascending slot)."""
        t0 = time.perf_counter()
        idx = np.asarray(idx)
        idx = idx[np.argsort(self.slots[idx])]
        act = self.slots[idx].astype(np.int32)
        if not self._rows_ready:
            tc = _encode_treecodes(self.dim, self.level, self.pos, self.forest.Jmax)
            self.sol.set_grid(self.slots.astype(np.int32), self._lvl32, tc, act)
            self._rows_ready = True
            self._tick("wgpu_set_grid (tree)", t0)
        else:
            self.sol.set_active(act)
            self._tick("wgpu_set_active", t0)
        return idx

    # ------------------------------------------------------------------ wavelet_decompose_full_tree + coarseningIndicator_tree
    def decompose(self, eps: Optional[float] = None, norm=None, eps_norm: str = "Linfty", thresh_comp=None, threshold: bool = True,
                  want_dict: bool = True):
        sol, dim = self.sol, self.dim
        nd = 2 ** dim

        def flags(idx):
            if not threshold:
                return
            st, det = sol.threshold_tree(WD, eps=eps, norm=norm, eps_norm=eps_norm, thresh_comp=thresh_comp, level_ref=self.forest.Jmax,
                                         want_detail=True)
            if self.det is None:
                self.det = np.zeros((len(self.code), det.shape[1]))
            self.st[idx] = st
            self.det[idx] = det

        def d2m(level):
            if level <= self.Jmin:
                return
            m = np.flatnonzero((self.level == level - 1) & (self.child[:, 0] >= 0))
            if len(m) == 0:
                return
            da = np.zeros((len(m), nd), dtype=np.int32)
            for c in range(nd):          # treecode digit order: bit0 -> y, bit1 -> x, bit2 -> z
                col = ((c >> 1) & 1) + 2 * (c & 1) + (4 * ((c >> 2) & 1) if dim == 3 else 0)
                j = self.child[m, col]
                assert (j >= 0).all()
                da[:, c] = self.slots[j]
            sol.coarsen_blocks(self.slots[m].astype(np.int32), da.ravel(), WD)

        if self.leaf_first:
            sol.set_forest(self.forest)                                   # leaf grid: full synchronisation, filtered restriction
            self._rows_ready = False                                      # (registers the leaves only: the tree is re-registered below)
            leaves = np.flatnonzero(self.is_leaf)
            leaves = leaves[np.argsort(self.slots[leaves])]
            sol.waveletDecomposition_tree((HVY_BLOCK, 0), WD)
            if self.lifted:
                sol.coarse_extension_modify(WD, (HVY_BLOCK, 0), True, True)
            flags(leaves)
        for level in range(self.Jmax_active, self.Jmin - 1, -1):
            todo = np.flatnonzero((self.level == level) & ~(self.is_leaf if self.leaf_first else np.zeros(len(self.code), bool)))
            if len(todo):
                idx = self.set_pass_topology(todo)
                t0 = time.perf_counter()
                sol.waveletDecomposition_tree((HVY_BLOCK, 0), WD)
                if self.lifted and self.is_leaf[idx].any():
                    sol.coarse_extension_modify(WD, (HVY_BLOCK, 0), True, True)
                t0 = self._tick("fwt + ce kernels", t0)
                flags(idx)
                self._tick("threshold", t0)
            t0 = time.perf_counter()
            d2m(level)
            self._tick("d2m", t0)
        return self.status_dict(self.st) if want_dict else None

    def status_dict(self, st):
        return dict(zip(self.keys(), (int(v) for v in st)))

    # ------------------------------------------------------------------ grid decision (light data)
    def decide(self, st0: np.ndarray) -> np.ndarray:
        """respectJmaxJmin_tree + ensureGradedness_tree(check_daughters) on the full tree (LIB/MESH/ensureGradedness_tree.f90,
        ensure_completeness_block.f90): a block keeps -1 only if it sits above Jmin, all its sisters carry -1, none of its daughters stays
        and no finer neighbour stays.  Statuses only move from -1 to "stay" (9), so the fixed point does not depend on the sweep order
        (libwabbit_host.so: whost_ft_decide)."""
        st = np.ascontiguousarray(st0, dtype=np.int32).copy()
        i32 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        rc = host_lib().whost_ft_decide(self.dim, len(st), i32(self._lvl32), i32(self.nb), i32(self.par), i32(self.child), self.Jmin, i32(st))
        if rc:
            raise RuntimeError(f"whost_ft_decide: {rc}")
        return st

    def security_zone(self, st0: np.ndarray, eps, norm, thresh_comp=None, force_maxlevel_dealiasing: bool = False, eps_norm: str = "Linfty") -> np.ndarray:
        """addSecurityZone_CE_tree (LIB/MESH/securityZone_tree.f90:140-298): an insignificant block (-1) next to a significant same-level block
        stays (0) if the significant block has significant details inside the Nwc-deep strip at their interface -- details the coarse
        extension would delete if the neighbour were coarsened.  The pairs are evaluated on the device (wgpu_patch_details_norm: coefficients
        renormalised for eps_norm as threshold_block does)."""
        sol, dim = self.sol, self.dim
        nc = sol.params.n_eqn
        eps = sol.params.eps if eps is None else eps
        st = st0.copy()
        sig = st0 == 0
        if force_maxlevel_dealiasing:
            sig &= self.level != self.forest.Jmax
        i32 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        t0 = time.perf_counter()
        sig8 = np.ascontiguousarray(sig, dtype=np.uint8)
        st32 = np.ascontiguousarray(st0, dtype=np.int32)
        cap = max(4 * int(sig8.sum()) + 64, 1024)
        while True:
            b, q = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
            k = host_lib().whost_ft_security_pairs(dim, len(st32), i32(self.nb), i32(st32), sig8.ctypes.data_as(C.POINTER(C.c_uint8)), cap, i32(b), i32(q))
            if k == -2:
                cap *= 4
                continue
            if k < 0:
                raise RuntimeError("whost_ft_security_pairs failed")
            b, q = b[:k].astype(np.int64), q[:k].astype(np.int64)
            break
        if len(b) == 0:
            return st
        dcode = np.array([(d[2] + 1) * 9 + (d[1] + 1) * 3 + (d[0] + 1) for d in self.dirs], dtype=np.int32)
        if getattr(self, "timing", None) is not None:
            t0 = self._tick(f"  sz: pairs (host)", t0)
        det = sol.patch_details(self.slots[b].astype(np.int32), dcode[q], WD, eps_norm=eps_norm, level_ref=self.forest.Jmax)
        if getattr(self, "timing", None) is not None:
            self.timing["  sz: n_pairs"] = len(b)
            t0 = self._tick("  sz: wgpu_patch_details", t0)
        tc = np.ones(nc, np.int32) if thresh_comp is None else np.asarray(thresh_comp, dtype=np.int32)
        e = np.full(nc, eps, dtype=np.float64) * (1.0 if norm is None else np.asarray(norm, dtype=np.float64))
        d_use = np.where(tc[None, :] == 0, 0.0, det)
        for l in range(2, int(tc.max()) + 1):                    # joint groups: max over the group's components (threshold_block.f90:96-99)
            grp = tc == l
            if grp.any():
                d_use[:, grp] = det[:, grp].max(axis=1, keepdims=True)
        significant = (d_use > e[None, :]).any(axis=1)
        st[self.nb[b[significant], q[significant]]] = 0
        return st

    def _with_marked_neighbours(self, marked: np.ndarray) -> np.ndarray:
        """Bs < Nrecon (adapt_tree.f90:771-803, reconstruct_neighbors): the leaves with a same-level neighbour that is marked for reconstruction
        are reconstructed as well (the modified coefficients of an interface block reach into their reconstruction)"""
        is_marked = np.zeros(len(self.code), bool)
        is_marked[marked] = True
        leaves = np.flatnonzero(self.is_leaf & ~is_marked)
        nb = self.nb[leaves]
        hit = (np.where(nb >= 0, is_marked[np.maximum(nb, 0)], False)).any(axis=1)
        return np.sort(np.concatenate([marked, leaves[hit]]))

    def _ce_sizes(self):
        """Nrecon and Ndep2 of setup_wavelet incl. the widening to the FD stencil (module_wavelets.f90:1368-1417)"""
        p = self.sol.params
        w = p.wavelet
        X, Y = int(w[3]), int(w[4])
        F = (X - 1) + (Y - 1) if Y > 0 else 0          # half width of HD (and of GR)
        H = {"FD_2nd_central": 1, "FD_4th_central": 2, "FD_6th_central": 3, "FD_4th_central_optimized": 3}[p.discretization]
        nwl, nwr = max(F - 1, 0) + (X - 1), F + (X - 1)
        dl, dr = max(2 * H - nwl, 0), max(2 * H - nwr, 0)
        nrl, nrr = nwl + F + dl, nwr + F + dr
        return nrl, nrr, nrl + max(X // 2 - 1, 0), nrr + X // 2

    # ------------------------------------------------------------------ adapt_tree
    def adapt(self, eps: Optional[float] = None, norm=None, eps_norm: str = "Linfty", thresh_comp=None, force_maxlevel_dealiasing: bool = False,
              indicator: str = "threshold-state-vector", want_info: bool = True, use_security_zone: bool = False, mask_keeps=None):
        """adapt_tree (LIB/MESH/adapt_tree.f90:11-260) for a lifted wavelet with the coarse extension, optionally the security zone: full-tree
        decomposition and indicator, grid decision, coarse extension on the lasting coarse/fine interfaces, reconstruction of the leaves at
        those interfaces (all at once if Bs >= Ndep2, else level by level from coarse to fine), pruning to the leaves, blocks moved to
        their places along the space-filling curve.  Returns (new forest, info)."""
        sol, dim = self.sol, self.dim
        p = sol.params
        if indicator == "everywhere":
            self.decompose(threshold=False, want_dict=False)
            st0 = np.where(self.is_leaf, -1, 0).astype(np.int32)
        else:
            self.decompose(eps=eps, norm=norm, eps_norm=eps_norm, thresh_comp=thresh_comp, want_dict=False)
            st0 = self.st.copy()
            if force_maxlevel_dealiasing:
                st0[self.level == self.forest.Jmax] = -1
        if mask_keeps is not None and indicator != "everywhere":
            # threshold_mask (coarseningIndicatorMask_tree, coarseningIndicator_tree.f90:290-331): blocks of the tree whose mask function is
            # not constant stay; mask_keeps(level, pos) is host geometry code
            cand = st0 == -1
            if force_maxlevel_dealiasing:
                cand &= self.level != self.forest.Jmax
            ci = np.flatnonzero(cand)
            if len(ci):
                st0[ci[np.asarray(mask_keeps(self.level[ci], self.pos[ci]), dtype=bool)]] = 0
        t0 = time.perf_counter()
        if use_security_zone and indicator != "everywhere":
            st0 = self.security_zone(st0, eps, norm, thresh_comp, force_maxlevel_dealiasing, eps_norm)
            t0 = self._tick("security zone", t0)
        st = self.decide(st0)
        t0 = self._tick("decide", t0)
        info = {"status0": self.status_dict(st0), "status": self.status_dict(st)} if want_info else {}
        keep = st != -1
        st_kept = st[keep]
        self.code, self.level, self.pos, self.slots = self.code[keep], self.level[keep], self.pos[keep], self.slots[keep]
        self._build_tables()
        t0 = self._tick("rebuild tables", t0)
        self.is_leaf = self.child[:, 0] < 0
        leaves = np.flatnonzero(self.is_leaf)
        marked = leaves[(self.nb[leaves] < 0).any(axis=1)]
        nrl, nrr, d2l, d2r = self._ce_sizes()
        leaf_only = all(p.Bs[a] >= d2l and p.Bs[a] >= d2r for a in range(dim))
        if not self.lifted:
            marked = marked[:0]      # no coarse extension: every block keeps its original / assembled values (adapt_tree.f90:236-241)
        elif any(p.Bs[a] < max(nrl, nrr) for a in range(dim)):
            marked = self._with_marked_neighbours(marked)
        if len(marked):
            self.set_pass_topology(marked)                                # the lasting interfaces (adapt_tree.f90:222-228): same-level
            sol.coarse_extension_modify(WD, (HVY_BLOCK, 0), True, False)  # neighbours send coefficients that carry the extension
            if leaf_only:
                sol.waveletReconstruction_CE(WD, (HVY_BLOCK, 0), (HVY_BLOCK, 0))
            else:
                for level in range(self.Jmin, int(self.level.max()) + 1):
                    todo = marked[self.level[marked] == level]
                    if len(todo):
                        self.set_pass_topology(todo)
                        sol.waveletReconstruction_CE(WD, (HVY_BLOCK, 0), (HVY_BLOCK, 0))
        # prune_fulltree2leafs + balanceLoad_tree: the leaves move to their slots along the space-filling curve
        new = Forest.from_blocks(dim, self.forest.Jmax, self.level[leaves].astype(np.int32), self.pos[leaves].astype(np.int32),
                                 block_dist=self.forest.block_dist, n_ranks=1, max_blocks=self.forest.max_blocks, periodic=self.forest.periodic)
        hvy, lvl, ixyz, _ = new.active(0)
        at = self._find(lvl.astype(np.int64), ixyz.astype(np.int64))
        src = self.slots[at]
        # lgt_block(:, IDX_REFINE_STS) of the new leaves: 0 significant / 9 REF_UNSIGNIFICANT_STAY, read by the "significant" refinement indicator
        self.leaf_status = st_kept[at].astype(np.int32)
        t0 = self._tick("reconstruction passes + new forest", t0)
        sol.move_blocks(src.astype(np.int32), hvy.astype(np.int32))
        sol.set_forest(new)
        self._tick("move_blocks + set_forest", t0)
        if self.timing is not None:
            print("FullTree.adapt timing [ms]:", {k: round(v * 1e3, 1) for k, v in self.timing.items()}, flush=True)
        if want_info:
            info.update({"marked": sorted(self.keys(marked)), "leaf_only": leaf_only, "leaf_first": self.leaf_first})
        return new, info


# ----------------------------------------------------------------------------------------------------------------------
# The same algorithm with the blocks partitioned over several GPUs
# ----------------------------------------------------------------------------------------------------------------------
class DistributedFullTree(FullTree):
    """adapt_tree with the full wavelet transformation across ranks.  The tree (light data) is replicated; every leaf lives on its owner
    in the forest's partition, every mother on the owner of its first daughter, in a free slot behind that rank's leaves.  A pass runs on
    every rank for the blocks it owns; the same-level neighbours (and, for the reconstruction, the coarser leaves) that other ranks own are
    shipped as whole blocks into scratch slots first (block_xfer: wgpu_gather_blocks -> all-to-all -> wgpu_scatter_blocks), so the kernels
    never know about ranks.  Flags are all-gathered (synchronize_lgt_data); the grid decision is computed on every rank."""

    def __init__(self, drv, Jmin: int = 1):
        # drv: wabbit_b200.multi.DistributedWabbit (sol, forest, rank, world, transport, _ship)
        from .multi import tick
        self._tk = tick
        _t_init = tick("adapt: tree init (numpy + tables)").__enter__()
        self.drv, self.me, self.world = drv, drv.rank, drv.world
        sol, forest = drv.sol, drv.forest
        self.sol, self.forest, self.dim, self.Jmin = sol, forest, forest.dim, Jmin
        dim = self.dim
        lv, ix, ow, sl = [], [], [], []
        for r in range(self.world):
            hvy, lvl, ixyz, _ = forest.active(r)
            lv.append(lvl.astype(np.int64))
            ix.append(ixyz.astype(np.int64))
            ow.append(np.full(len(hvy), r, np.int64))
            sl.append(hvy.astype(np.int64))
        l0, p0, o0, s0 = np.concatenate(lv), np.concatenate(ix), np.concatenate(ow), np.concatenate(sl)
        level, pos, leaf_of = _build_tree(dim, Jmin, l0, p0)
        is_leaf = leaf_of >= 0
        owner = np.full(len(level), -1, np.int64)
        slots = np.zeros(len(level), np.int64)
        owner[is_leaf], slots[is_leaf] = o0[leaf_of[is_leaf]], s0[leaf_of[is_leaf]]
        self._set_blocks(level, pos, slots, is_leaf, presorted=True)
        self.owner = owner
        # mothers: owner = owner of the first daughter, finest level first; slots behind the rank's leaves, in tree order
        nxt = np.array([forest.n_active(r) + 1 for r in range(self.world)], dtype=np.int64)
        for L in range(int(self.level.max()) - 1, Jmin - 1, -1):
            m = np.flatnonzero((self.level == L) & ~self.is_leaf)
            self.owner[m] = self.owner[self.child[m, 0]]
            for r in range(self.world):
                mr = m[self.owner[m] == r]
                self.slots[mr] = nxt[r] + np.arange(len(mr))
                nxt[r] += len(mr)
        self.scratch0 = nxt                                       # first free slot per rank behind leaves and mothers
        if nxt.max() - 1 > sol.max_blocks:
            raise MemoryError(f"full tree needs {int(nxt.max()) - 1} block slots on a rank, max_blocks = {sol.max_blocks}")
        self.Jmax_active = int(self.level[self.is_leaf].max())
        self.st = np.zeros(len(self.level), np.int32)
        self.det = None
        self.timing = None
        F = sol.wavelet_filter_width()
        p = sol.params
        self.leaf_first = all(p.Bs[a] >= 3 * F for a in range(dim))
        self.lifted = (p.wavelet[4] != "0") if p.useCoarseExtension < 0 else bool(p.useCoarseExtension)
        self._halo_cleared = False
        _t_init.__exit__()

    # ------------------------------------------------------------------ a pass: ship what the owned blocks need, then local tables
    def _pass(self, active: np.ndarray, need):
        """active: tree indices of the blocks of the pass (all ranks).  need(idx_of_rank_q) -> list of (array, tree indices of the source
        blocks rank q reads).  Ships the remote ones, uploads the local topology, returns my active indices in active-list order."""
        sol, drv, me, W, dim = self.sol, self.drv, self.me, self.world, self.dim
        if not self._halo_cleared:      # the time stepper's halo slots are dead now; their slots are reused by mothers and copies
            z = np.zeros(0, np.int32)
            sol._check(sol._lib.wgpu_set_halo(sol._ctx, 0, z.ctypes.data_as(C.POINTER(C.c_int32)), z.ctypes.data_as(C.POINTER(C.c_int32)),
                                              z.ctypes.data_as(C.POINTER(C.c_int32)), 0, z.ctypes.data_as(C.POINTER(C.c_int32)), None))
            self._halo_cleared = True
        _t = self._tk("adapt: pass lists (numpy)").__enter__()
        mine_of = [active[self.owner[active] == q] for q in range(W)]
        lslot = np.where(self.owner == me, self.slots, -1)                  # local slot of every tree block present on this rank
        nxt = int(self.scratch0[me])
        needs = [need(mine_of[q]) for q in range(W)]
        n_arrays = len(needs[0])
        _t.__exit__()
        # a block may be read from several arrays in one pass (coefficients as the same-level neighbour of one block, values as the coarser
        # neighbour of another): it has ONE local slot, so every array of the pass ships the union of the blocks any of them needs
        src_r, src_s, dst_r, which = [], [], [], []
        seen = np.zeros(len(self.code), bool)
        for q in range(W):
            b = np.concatenate([needs[q][a][1] for a in range(n_arrays)]) if n_arrays else np.zeros(0, np.int64)
            b = b[b >= 0]
            b = b[self.owner[b] != q]
            seen[:] = False                      # sorted unique by a mark array: no sort of the 26-neighbour lists
            seen[b] = True
            b = np.flatnonzero(seen)
            src_r.append(self.owner[b])
            src_s.append(self.slots[b])
            dst_r.append(np.full(len(b), q, np.int64))
            which.append(b)
        src_r, src_s, dst_r, which = (np.concatenate(v) for v in (src_r, src_s, dst_r, which))
        first = nxt
        for a in range(n_arrays):
            with self._tk("adapt: pass ship", sol):
                loc, nxt = drv._ship(needs[0][a][0], src_r, src_s, dst_r, first)
            lslot[which[dst_r == me]] = loc
        mine = mine_of[me]
        mine = mine[np.argsort(self.slots[mine])]
        # every block of the tree is registered by position (resident ones with their local slot, the others with slot 0 = "exists, no data
        # here"); the relations of the pass's blocks are derived on the device
        if getattr(self, "_tc_all", None) is None or len(self._tc_all) != len(self.code):
            self._tc_all = _encode_treecodes(dim, self.level, self.pos, self.forest.Jmax)
        with self._tk("adapt: pass set_grid", sol):
            sol.set_grid(np.maximum(lslot, 0).astype(np.int32), self._lvl32, self._tc_all, self.slots[mine].astype(np.int32))
        self._lslot = lslot
        return mine, mine_of

    def _gather_status(self, mine_of, st_local, det_local=None):
        counts = [len(m) for m in mine_of]
        st = self.drv.tr.allgather_np(np.asarray(st_local, dtype=np.int32), counts)
        off = 0
        for q in range(self.world):
            m = mine_of[q][np.argsort(self.slots[mine_of[q]])]
            self.st[m] = st[off:off + len(m)]
            off += len(m)

    # ------------------------------------------------------------------ decomposition + indicator
    def decompose(self, eps=None, norm=None, eps_norm: str = "Linfty", thresh_comp=None, threshold: bool = True, want_dict: bool = False):
        sol, drv, dim, me = self.sol, self.drv, self.dim, self.me
        nd = 2 ** dim
        same_level = lambda idx: [((HVY_BLOCK, 0), self.nb[idx].ravel())]

        def flags(mine, mine_of):
            if not threshold:
                return
            with self._tk("adapt: threshold + allgather", sol):
                st = sol.threshold_tree(WD, eps=eps, norm=norm, eps_norm=eps_norm, thresh_comp=thresh_comp, level_ref=self.forest.Jmax) if len(mine) \
                    else np.zeros(0, np.int32)
                self._gather_status(mine_of, st)

        def d2m(level):
            if level <= self.Jmin:
                return
            m_all = np.flatnonzero((self.level == level - 1) & (self.child[:, 0] >= 0))
            if len(m_all) == 0:
                return
            # sync_D2M across ranks: the decomposed daughters travel to the owner of their mother
            cols = [((c >> 1) & 1) + 2 * (c & 1) + (4 * ((c >> 2) & 1) if dim == 3 else 0) for c in range(nd)]   # treecode digit order
            da = self.child[m_all][:, cols]                                   # [n_m, nd] tree indices
            src_r, src_s = self.owner[da.ravel()], self.slots[da.ravel()]
            dst_r = np.repeat(self.owner[m_all], nd)
            with self._tk("adapt: d2m ship + coarsen", sol):
                loc, _ = drv._ship(WD, src_r, src_s, dst_r, int(self.scratch0[me]))
                mine = m_all[self.owner[m_all] == me]
                if len(mine):
                    sol.coarsen_blocks(self.slots[mine].astype(np.int32), loc.astype(np.int32), WD)

        if self.leaf_first:
            # leaf pass on the time stepper's topology: halo copies (and filtered copies of finer neighbours) of hvy_block refreshed first
            with self._tk("adapt: leaf pass (exchange + fwt + ce)", sol):
                drv.stepper.exchange_array(0, 0)
                sol.waveletDecomposition_tree((HVY_BLOCK, 0), WD)
                if self.lifted:
                    sol.coarse_extension_modify(WD, (HVY_BLOCK, 0), True, True)
            leaves = np.flatnonzero(self.is_leaf)
            mine_of = [leaves[self.owner[leaves] == q] for q in range(self.world)]
            if threshold:
                with self._tk("adapt: threshold + allgather", sol):
                    st = sol.threshold_tree(WD, eps=eps, norm=norm, eps_norm=eps_norm, thresh_comp=thresh_comp, level_ref=self.forest.Jmax)
                    self._gather_status(mine_of, st)
        for level in range(self.Jmax_active, self.Jmin - 1, -1):
            todo = np.flatnonzero((self.level == level) & ~(self.is_leaf if self.leaf_first else np.zeros(len(self.code), bool)))
            if len(todo):
                mine, mine_of = self._pass(todo, same_level)
                if len(mine):
                    with self._tk("adapt: fwt + ce kernels", sol):
                        sol.waveletDecomposition_tree((HVY_BLOCK, 0), WD)
                        if self.lifted and self.is_leaf[mine].any():
                            sol.coarse_extension_modify(WD, (HVY_BLOCK, 0), True, True)
                flags(mine, mine_of)
            d2m(level)
        return None

    def security_zone(self, st0, eps, norm, thresh_comp=None, force_maxlevel_dealiasing: bool = False, eps_norm: str = "Linfty"):
        """addSecurityZone_CE_tree across ranks: every rank evaluates the strips of the significant blocks it owns; the kept neighbours are
        merged over the ranks (synchronize_lgt_data)."""
        me = self.me
        mask = self.owner == me
        local = FullTree.security_zone(self, np.where(mask | (st0 != 0), st0, 1).astype(np.int32), eps, norm, thresh_comp,
                                       force_maxlevel_dealiasing, eps_norm)        # blocks of other ranks are not "significant" here (status 1)
        kept = (st0 == -1) & (local == 0)
        allk = self.drv.tr.allreduce_max_np(kept.astype(np.float64))
        st = st0.copy()
        st[allk > 0] = 0
        return st

    # ------------------------------------------------------------------ adapt_tree
    def adapt(self, eps=None, norm=None, eps_norm: str = "Linfty", thresh_comp=None, force_maxlevel_dealiasing: bool = False,
              indicator: str = "threshold-state-vector", want_info: bool = False, use_security_zone: bool = False, mask_keeps=None):
        sol, drv, dim, me, W = self.sol, self.drv, self.dim, self.me, self.world
        p = sol.params
        if indicator == "everywhere":
            self.decompose(threshold=False)
            st0 = np.where(self.is_leaf, -1, 0).astype(np.int32)
        else:
            self.decompose(eps=eps, norm=norm, eps_norm=eps_norm, thresh_comp=thresh_comp)
            st0 = self.st.copy()
            if force_maxlevel_dealiasing:
                st0[self.level == self.forest.Jmax] = -1
        if mask_keeps is not None and indicator != "everywhere":      # threshold_mask: replicated light data, as on one rank
            cand = st0 == -1
            if force_maxlevel_dealiasing:
                cand &= self.level != self.forest.Jmax
            ci = np.flatnonzero(cand)
            if len(ci):
                st0[ci[np.asarray(mask_keeps(self.level[ci], self.pos[ci]), dtype=bool)]] = 0
        if use_security_zone and indicator != "everywhere":
            self._lslot_for_patches()
            with self._tk("adapt: security zone", sol):
                st0 = self.security_zone(st0, eps, norm, thresh_comp, force_maxlevel_dealiasing, eps_norm)
        with self._tk("adapt: decide"):
            st = self.decide(st0)
        _t = self._tk("adapt: prune + tables").__enter__()
        keep = st != -1
        st_kept = st[keep]
        self.code, self.level, self.pos, self.slots, self.owner = (a[keep] for a in (self.code, self.level, self.pos, self.slots, self.owner))
        self._build_tables()
        _t.__exit__()
        self.is_leaf = self.child[:, 0] < 0
        leaves = np.flatnonzero(self.is_leaf)
        marked = leaves[(self.nb[leaves] < 0).any(axis=1)] if self.lifted else leaves[:0]
        nrl, nrr, d2l, d2r = self._ce_sizes()
        if self.lifted and any(p.Bs[a] < max(nrl, nrr) for a in range(dim)):
            marked = self._with_marked_neighbours(marked)
        leaf_only = all(p.Bs[a] >= d2l and p.Bs[a] >= d2r for a in range(dim))
        _t_rec = self._tk("adapt: CE + reconstruction passes", sol).__enter__()
        if len(marked):
            nothing = lambda idx: [((HVY_BLOCK, 0), np.zeros(0, np.int64))]
            mine, _ = self._pass(marked, nothing)                              # coarse extension on the lasting interfaces (local data only)
            if len(mine):
                sol.coarse_extension_modify(WD, (HVY_BLOCK, 0), True, False)

            def recon_needs(idx):
                # coefficients of the same-level neighbours; values of the coarser leaves that cover the directions without one
                coarse = []
                for q, d in enumerate(self.dirs):
                    miss = idx[self.nb[idx, q] < 0]
                    if len(miss):
                        coarse.append(self._find(self.level[miss] - 1, self._neighbor_pos(miss, d) >> 1))
                return [(WD, self.nb[idx].ravel()), ((HVY_BLOCK, 0), np.concatenate(coarse) if coarse else np.zeros(0, np.int64))]

            levels = [None] if leaf_only else list(range(self.Jmin, int(self.level.max()) + 1))
            for level in levels:
                todo = marked if level is None else marked[self.level[marked] == level]
                if len(todo):
                    mine, _ = self._pass(todo, recon_needs)
                    if len(mine):
                        sol.waveletReconstruction_CE(WD, (HVY_BLOCK, 0), (HVY_BLOCK, 0))
        _t_rec.__exit__()
        # prune_fulltree2leafs + balanceLoad_tree: the leaves move to their owners / slots in the new partition
        _t = self._tk("adapt: new forest + lists").__enter__()
        new = Forest.from_blocks(dim, self.forest.Jmax, self.level[leaves].astype(np.int32), self.pos[leaves].astype(np.int32),
                                 block_dist=self.forest.block_dist, n_ranks=W, max_blocks=self.forest.max_blocks, periodic=self.forest.periodic)
        src_r, src_s, dst_r, dst_s, stat = [], [], [], [], []
        for r in range(W):
            hvy, lvl, ixyz, _ = new.active(r)
            i = self._find(lvl.astype(np.int64), ixyz.astype(np.int64))
            stat.append(st_kept[i])
            src_r.append(self.owner[i])
            src_s.append(self.slots[i])
            dst_r.append(np.full(len(hvy), r, np.int64))
            dst_s.append(hvy.astype(np.int64))
        src_r, src_s, dst_r, dst_s = (np.concatenate(v) for v in (src_r, src_s, dst_r, dst_s))
        self.leaf_status = np.concatenate(stat).astype(np.int32)      # refinement status of the new leaves, global space-filling-curve order
        _t.__exit__()
        with self._tk("adapt: final ship + move", sol):
            loc, _ = drv._ship((HVY_BLOCK, 0), src_r, src_s, dst_r, int(self.scratch0[me]))
            sol.move_blocks(loc.astype(np.int32), dst_s[dst_r == me].astype(np.int32))
        with self._tk("adapt: attach (halo plan + set_grid)", sol):
            drv.attach(new)
        return new, {}

    def _lslot_for_patches(self):
        """wgpu_patch_details addresses blocks by slot: own blocks only (tree slots are the owner's)"""
        return None
