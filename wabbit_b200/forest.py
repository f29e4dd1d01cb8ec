"""Host forest metadata (light data) -- thin wrapper over libwabbit_host.so (include/wabbit_host.h).

Stands in for what WABBIT's host Fortran (createEquidistantGrid_tree, updateMetadata_tree,
balanceLoad_tree) leaves in lgt_block / hvy_active / hvy_neighbor.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from ._native import host_lib

SFC = {"sfc_z": 0, "sfc_hilbert": 1}


def _i32(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class _ForestView(np.ndarray):
    """read-only view of a forest's neighbour table that keeps the forest alive"""

    def __new__(cls, arr, owner):
        obj = np.asarray(arr).view(cls)
        obj._owner = owner
        return obj

    def __array_finalize__(self, obj):
        self._owner = getattr(obj, "_owner", None)


class Forest:
    def __init__(self, handle, dim: int, Jmax: int, n_ranks: int, max_blocks: int, block_dist: str = "sfc_hilbert",
                 periodic: Sequence[int] = (1, 1, 1)):
        self._h = handle
        self.dim, self.Jmax, self.n_ranks, self.max_blocks = dim, Jmax, n_ranks, max_blocks
        self.block_dist, self.periodic = block_dist, tuple(int(v) for v in periodic)

    @classmethod
    def uniform(cls, dim: int, J: int, Jmax: Optional[int] = None, block_dist: str = "sfc_hilbert", n_ranks: int = 1,
                max_blocks: Optional[int] = None, periodic: Sequence[int] = (1, 1, 1)) -> "Forest":
        Jmax = J if Jmax is None else Jmax
        nb = (2 ** J) ** dim
        max_blocks = max_blocks or -(-nb // n_ranks)
        h = C.c_void_p()
        per = np.asarray(periodic, dtype=np.int32)
        rc = host_lib().whost_create_uniform(dim, J, Jmax, SFC[block_dist], n_ranks, max_blocks, _i32(per), C.byref(h))
        if rc:
            raise RuntimeError(f"whost_create_uniform failed with code {rc}")
        return cls(h, dim, Jmax, n_ranks, max_blocks, block_dist, periodic)

    @classmethod
    def from_blocks(cls, dim: int, Jmax: int, level, ixyz, block_dist: str = "sfc_hilbert", n_ranks: int = 1,
                    max_blocks: Optional[int] = None, periodic: Sequence[int] = (1, 1, 1)) -> "Forest":
        level = np.ascontiguousarray(level, dtype=np.int32)
        ixyz = np.ascontiguousarray(ixyz, dtype=np.int32).reshape(-1, 3)
        n = len(level)
        max_blocks = max_blocks or -(-n // n_ranks)
        h = C.c_void_p()
        per = np.asarray(periodic, dtype=np.int32)
        rc = host_lib().whost_create_from_blocks(dim, Jmax, SFC[block_dist], n_ranks, max_blocks, _i32(per), n, _i32(level),
                                                 _i32(ixyz), C.byref(h))
        if rc:
            raise RuntimeError(f"whost_create_from_blocks failed with code {rc}")
        return cls(h, dim, Jmax, n_ranks, max_blocks, block_dist, periodic)

    # ------------------------------------------------------------------ grid adaptation (light data, single rank)
    def refine(self, flags=None, max_blocks: Optional[int] = None):
        """New forest with the flagged blocks (None: all below Jmax) replaced by their 2^dim daughters, ordered along the SFC, plus
        the id lists wgpu_refine consumes: (new_forest, mothers, daughters, keep_src, keep_dst), all 1-based hvy ids."""
        n, nd = self.n_active(0), 2 ** self.dim
        N = max_blocks or self.max_blocks
        fl = None if flags is None else np.ascontiguousarray(flags, dtype=np.int32)
        mo, da = np.zeros(n, np.int32), np.zeros(n * nd, np.int32)
        ks, kd = np.zeros(n, np.int32), np.zeros(n, np.int32)
        nm, nk = C.c_int32(), C.c_int32()
        h = C.c_void_p()
        rc = host_lib().whost_refine(self._h, None if fl is None else _i32(fl), N, C.byref(h), C.byref(nm), _i32(mo), _i32(da), C.byref(nk),
                                     _i32(ks), _i32(kd))
        if rc:
            raise MemoryError("refine: the refined grid needs more than max_blocks blocks") if rc == 2 else RuntimeError(f"whost_refine: {rc}")
        new = Forest(h, self.dim, self.Jmax, 1, N, self.block_dist, self.periodic)
        return new, mo[:nm.value], da[:nm.value * nd], ks[:nk.value], kd[:nk.value]

    def coarsen(self, status, Jmin: int = 1, max_blocks: Optional[int] = None):
        """Final refinement status (completeness, gradedness, Jmin) and the coarsened forest with the id lists wgpu_move_blocks /
        wgpu_coarsen consume: (new_forest, final_status, mothers[new ids], daughters[old ids], keep_src, keep_dst)."""
        n, nd = self.n_active(0), 2 ** self.dim
        N = max_blocks or self.max_blocks
        st = np.ascontiguousarray(status, dtype=np.int32).copy()
        mo, da = np.zeros(n, np.int32), np.zeros(n, np.int32)
        ks, kd = np.zeros(n, np.int32), np.zeros(n, np.int32)
        nm, nk = C.c_int32(), C.c_int32()
        h = C.c_void_p()
        rc = host_lib().whost_coarsen(self._h, _i32(st), Jmin, N, C.byref(h), C.byref(nm), _i32(mo), _i32(da), C.byref(nk), _i32(ks), _i32(kd))
        if rc:
            raise RuntimeError(f"whost_coarsen: {rc}")
        new = Forest(h, self.dim, self.Jmax, 1, N, self.block_dist, self.periodic)
        return new, st, mo[:nm.value], da[:nm.value * nd], ks[:nk.value], kd[:nk.value]

    def refine_global(self, flags=None):
        """refine() for a grid partitioned over any number of ranks: flags and all returned ids are 1-based positions in the global
        space-filling-curve order (rank-major order of the active lists); the new forest keeps n_ranks and max_blocks."""
        n, nd = self.n_blocks, 2 ** self.dim
        fl = None if flags is None else np.ascontiguousarray(flags, dtype=np.int32)
        mo, da = np.zeros(n, np.int32), np.zeros(n * nd, np.int32)
        ks, kd = np.zeros(n, np.int32), np.zeros(n, np.int32)
        nm, nk = C.c_int32(), C.c_int32()
        h = C.c_void_p()
        rc = host_lib().whost_refine_global(self._h, None if fl is None else _i32(fl), self.max_blocks, C.byref(h), C.byref(nm), _i32(mo), _i32(da),
                                            C.byref(nk), _i32(ks), _i32(kd))
        if rc:
            raise MemoryError("refine: the refined grid needs more than max_blocks blocks per rank") if rc == 2 else RuntimeError(f"whost_refine_global: {rc}")
        new = Forest(h, self.dim, self.Jmax, self.n_ranks, self.max_blocks, self.block_dist, self.periodic)
        return new, mo[:nm.value], da[:nm.value * nd], ks[:nk.value], kd[:nk.value]

    def coarsen_global(self, status, Jmin: int = 1):
        """coarsen() for a grid partitioned over any number of ranks (global positions, see refine_global)."""
        n, nd = self.n_blocks, 2 ** self.dim
        st = np.ascontiguousarray(status, dtype=np.int32).copy()
        mo, da = np.zeros(n, np.int32), np.zeros(n, np.int32)
        ks, kd = np.zeros(n, np.int32), np.zeros(n, np.int32)
        nm, nk = C.c_int32(), C.c_int32()
        h = C.c_void_p()
        rc = host_lib().whost_coarsen_global(self._h, _i32(st), Jmin, self.max_blocks, C.byref(h), C.byref(nm), _i32(mo), _i32(da), C.byref(nk),
                                             _i32(ks), _i32(kd))
        if rc:
            raise RuntimeError(f"whost_coarsen_global: {rc}")
        new = Forest(h, self.dim, self.Jmax, self.n_ranks, self.max_blocks, self.block_dist, self.periodic)
        return new, st, mo[:nm.value], da[:nm.value * nd], ks[:nk.value], kd[:nk.value]

    def __del__(self):
        try:
            if self._h:
                host_lib().whost_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def n_blocks(self) -> int:
        return host_lib().whost_n_blocks(self._h)

    def n_active(self, rank: int = 0) -> int:
        return host_lib().whost_n_active(self._h, rank)

    @property
    def is_uniform(self) -> bool:
        return bool(host_lib().whost_is_uniform(self._h))

    def active(self, rank: int = 0):
        """(hvy_active[1-based], level, ixyz[n,3], treecode) of a rank in SFC order."""
        n = self.n_active(rank)
        hvy = np.zeros(n, np.int32)
        lvl = np.zeros(n, np.int32)
        ixyz = np.zeros((n, 3), np.int32)
        tc = np.zeros(n, np.int64)
        host_lib().whost_get_active(self._h, rank, _i32(hvy), _i32(lvl), _i32(ixyz), tc.ctypes.data_as(C.POINTER(C.c_int64)))
        return hvy, lvl, ixyz, tc

    def neighbors(self, rank: int = 0) -> np.ndarray:
        """hvy_neighbor as an array [168, n_active(rank)] (== Fortran hvy_neighbor(1:hvy_n,168)), lgt ids, -1 none."""
        n = self.n_active(rank)
        if n == 0:
            return np.empty((168, 0), np.int32)
        ptr = host_lib().whost_neighbors_ptr(self._h, rank)
        out = np.ctypeslib.as_array(ptr, shape=(168, n))     # a view of the forest's own table (31 MB at 47k blocks): no copy
        out.flags.writeable = False
        return _ForestView(out, self)


def coarsening_groups(forest: "Forest", status: np.ndarray, Jmin: int = 1):
    """Host-side adapt logic for ONE coarsening sweep on a leaf grid (stand-in for respectJmaxJmin_tree, ensureCompleteness
    and ensureGradedness_tree, LIB/MESH/respectJmaxJmin_tree.f90, ensureGradedness_tree.f90:13): returns the final
    refinement status per active block -- -1 only for blocks whose 2^dim sisters all carry -1, sit above Jmin, and whose
    mother would not end up two levels coarser than any neighbour that stays.  Light data only; no heavy data is touched."""
    hvy, lvl, ixyz, _ = forest.active(0)
    n, dim, nd = len(hvy), forest.dim, 2 ** forest.dim
    st = np.where((np.asarray(status) == -1) & (lvl > Jmin), -1, 0).astype(np.int32)
    nbr = forest.neighbors(0)[:, :n]                      # [168, n], 1-based ids (single rank: lgt id == hvy id)
    pos = {int(h): k for k, h in enumerate(hvy)}
    key = lambda k: (int(lvl[k]) - 1, int(ixyz[k, 0]) // 2, int(ixyz[k, 1]) // 2, int(ixyz[k, 2]) // 2)
    changed = True
    while changed:
        changed = False
        groups = {}
        for k in range(n):
            if st[k] == -1:
                groups.setdefault(key(k), []).append(k)
        for m, ks in groups.items():
            ok = len(ks) == nd                            # completeness: all sisters are leaves and want to coarsen
            if ok:
                for k in ks:                              # gradedness: a finer neighbour must itself coarsen (to this level)
                    for slot in range(112, 168):
                        j = nbr[slot, k]
                        if j >= 1 and st[pos[int(j)]] != -1:
                            ok = False
                            break
                    if not ok:
                        break
            if not ok:
                for k in ks:
                    st[k] = 0
                changed = True
    return st
