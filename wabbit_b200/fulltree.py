"""adapt_tree with the full wavelet transformation for LIFTED wavelets, on the device (single rank).

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


class FullTree:
    """Leaves of `forest` (slots = their hvy ids) plus all ancestors down to Jmin in free slots behind them (init_full_tree)."""

    def __init__(self, sol, forest: Forest, Jmin: int = 1):
        self.sol, self.forest, self.dim, self.Jmin = sol, forest, forest.dim, Jmin
        hvy, lvl, ixyz, _ = forest.active(0)
        self.slot: Dict[Key, int] = {}
        self.leaf = set()
        for h, l, x in zip(hvy, lvl, ixyz):
            k = (int(l), int(x[0]), int(x[1]), int(x[2]))
            self.slot[k] = int(h)
            self.leaf.add(k)
        nxt = int(hvy.max()) + 1
        mothers = set()
        for k in self.leaf:
            while k[0] > Jmin:
                k = _parent(k)
                if k in mothers:
                    break
                mothers.add(k)
        for k in sorted(mothers):
            self.slot[k] = nxt
            nxt += 1
        if nxt - 1 > sol.max_blocks:
            raise MemoryError(f"full tree needs {nxt - 1} block slots, max_blocks = {sol.max_blocks}")
        self.Jmax_active = max(k[0] for k in self.leaf)
        self.status: Dict[Key, int] = {}
        self.detail: Dict[Key, np.ndarray] = {}
        F = sol.wavelet_filter_width()
        p = sol.params
        self.leaf_first = all(p.Bs[a] >= 3 * F for a in range(self.dim))

    # ------------------------------------------------------------------ per-pass topology
    def _nbr_key(self, k: Key, d) -> Key:
        n = 2 ** k[0]
        return (k[0],) + tuple(((k[1 + a] + d[a]) % n) if a < self.dim else 0 for a in range(3))

    def set_pass_topology(self, keys: List[Key]):
        """neighbour table of a block list: same-level relations to whatever block of the tree sits there (leaf or mother); for leaves,
        directions without a same-level block become coarser relations (slot + 56), which is where the coarse extension acts"""
        sol, dim = self.sol, self.dim
        lib = host_lib()
        ids = np.array([self.slot[k] for k in keys], dtype=np.int32)
        order = np.argsort(ids)
        keys = [keys[i] for i in order]
        ids = ids[order]
        ld = int(ids.max())
        nbr = np.full((168, ld), -1, dtype=np.int32)
        lvl = np.array([k[0] for k in keys], dtype=np.int32)
        ix = np.zeros(3, dtype=np.int32)

        def encode(k):
            ix[:] = k[1:]
            return lib.whost_encode(dim, k[0], self.forest.Jmax, ix.ctypes.data_as(C.POINTER(C.c_int32)))

        tc = np.array([encode(k) for k in keys], dtype=np.int64)
        coarse = set()                                  # coarser leaves next to the blocks of the pass: known as data sources only
        for i, k in enumerate(keys):
            s = self.slot[k] - 1
            for d in _dirs(dim):
                nk = self._nbr_key(k, d)
                if nk in self.slot:
                    nbr[_code(d) - 1, s] = self.slot[nk]
                elif k in self.leaf and k[0] > 0:
                    ck = _parent(nk)
                    if ck in self.slot:
                        nbr[_code(d) - 1 + 56, s] = self.slot[ck]
                        coarse.add(ck)
        coarse = sorted(coarse - set(keys))
        sol.set_treecodes(np.concatenate([ids, np.array([self.slot[k] for k in coarse], dtype=np.int32)]),
                          np.concatenate([lvl, np.array([k[0] for k in coarse], dtype=np.int32)]),
                          np.concatenate([tc, np.array([encode(k) for k in coarse], dtype=np.int64)]))
        sol.set_topology(ids, lvl, nbr, 0)
        return keys

    # ------------------------------------------------------------------ wavelet_decompose_full_tree + coarseningIndicator_tree
    def decompose(self, eps: Optional[float] = None, norm=None, eps_norm: str = "Linfty", thresh_comp=None, threshold: bool = True):
        sol, dim = self.sol, self.dim
        nd = 2 ** dim

        def flags(keys):
            if not threshold:
                return
            st, det = sol.threshold_tree(WD, eps=eps, norm=norm, eps_norm=eps_norm, thresh_comp=thresh_comp, level_ref=self.forest.Jmax,
                                         want_detail=True)
            for k, s, dd in zip(keys, st, det):
                self.status[k] = int(s)
                self.detail[k] = dd

        def d2m(level):
            ms = sorted({_parent(k) for k in self.slot if k[0] == level and level > self.Jmin} & set(self.slot))
            if not ms:
                return
            mo = np.array([self.slot[m] for m in ms], dtype=np.int32)
            da = np.array([self.slot[c] for m in ms for c in _children(m, dim)], dtype=np.int32)
            sol.coarsen_blocks(mo, da, WD)

        if self.leaf_first:
            sol.set_forest(self.forest)                                   # leaf grid: full synchronisation, filtered restriction
            keys = [k for _, k in sorted((self.slot[k], k) for k in self.leaf)]
            sol.waveletDecomposition_tree((HVY_BLOCK, 0), WD)
            sol.coarse_extension_modify(WD, (HVY_BLOCK, 0), True, True)
            flags(keys)
        for level in range(self.Jmax_active, self.Jmin - 1, -1):
            todo = [k for k in self.slot if k[0] == level and not (self.leaf_first and k in self.leaf)]
            if todo:
                keys = self.set_pass_topology(todo)
                sol.waveletDecomposition_tree((HVY_BLOCK, 0), WD)
                if any(k in self.leaf for k in keys):
                    sol.coarse_extension_modify(WD, (HVY_BLOCK, 0), True, True)
                flags(keys)
            d2m(level)
        return self.status


    # ------------------------------------------------------------------ grid decision (light data)
    def _finer_neighbors(self, k: Key):
        out = []
        for d in _dirs(self.dim):
            nk = self._nbr_key(k, d)
            for c in _children(nk, self.dim):
                if c in self.slot and all((d[a] == 0) or ((c[1 + a] & 1) == (0 if d[a] > 0 else 1)) for a in range(self.dim)):
                    out.append(c)
        return out

    def decide(self, st: Dict[Key, int]) -> Dict[Key, int]:
        """respectJmaxJmin_tree + ensureGradedness_tree(check_daughters) on the full tree (LIB/MESH/ensureGradedness_tree.f90,
        ensure_completeness_block.f90): a block keeps -1 only if it sits above Jmin, all its sisters carry -1, none of its daughters stays
        and no finer neighbour stays.  Statuses only move from -1 to "stay", so one monotone sweep to the fixed point."""
        st = dict(st)
        dim = self.dim
        for k in st:
            if st[k] == -1 and k[0] <= self.Jmin:
                st[k] = 9                                                  # REF_UNSIGNIFICANT_STAY
        changed = True
        while changed:
            changed = False
            for k in sorted(st):
                if st[k] != -1:
                    continue
                stay = any(st.get(s_, 0) != -1 for s_ in _children(_parent(k), dim))
                if not stay and k not in self.leaf:
                    stay = any(st.get(c, -1) != -1 for c in _children(k, dim) if c in self.slot)
                if not stay:
                    stay = any(st[f] != -1 for f in self._finer_neighbors(k))
                if stay:
                    st[k] = 9
                    changed = True
        return st

    def _ce_sizes(self):
        """Nrecon and Ndep2 of setup_wavelet incl. the widening to the FD stencil (module_wavelets.f90:1368-1417)"""
        p = self.sol.params
        w = p.wavelet
        X, Y = int(w[3]), int(w[4])
        F = (X - 1) + (Y - 1)
        H = {"FD_2nd_central": 1, "FD_4th_central": 2, "FD_6th_central": 3, "FD_4th_central_optimized": 3}[p.discretization]
        nwl, nwr = (F - 1) + (X - 1), F + (X - 1)
        dl, dr = max(2 * H - nwl, 0), max(2 * H - nwr, 0)
        nrl, nrr = nwl + F + dl, nwr + F + dr
        return nrl, nrr, nrl + max(X // 2 - 1, 0), nrr + X // 2

    # ------------------------------------------------------------------ adapt_tree
    def adapt(self, eps: Optional[float] = None, norm=None, eps_norm: str = "Linfty", thresh_comp=None, force_maxlevel_dealiasing: bool = False,
              indicator: str = "threshold-state-vector"):
        """adapt_tree (LIB/MESH/adapt_tree.f90:11-260) for a lifted wavelet with the coarse extension (useSecurityZone = 0): full-tree
        decomposition and indicator, grid decision, coarse extension on the lasting coarse/fine interfaces, reconstruction of the leaves at
        those interfaces (all at once if Bs >= Ndep2, else level by level from coarse to fine), pruning to the leaves, blocks moved to
        their places along the space-filling curve.  Returns (new forest, info)."""
        sol, dim = self.sol, self.dim
        p = sol.params
        if indicator == "everywhere":
            self.decompose(threshold=False)
            st0 = {k: (-1 if k in self.leaf else 0) for k in self.slot}
        else:
            st0 = dict(self.decompose(eps=eps, norm=norm, eps_norm=eps_norm, thresh_comp=thresh_comp))
            if force_maxlevel_dealiasing:
                st0 = {k: (-1 if k[0] == self.forest.Jmax else v) for k, v in st0.items()}
        st = self.decide(st0)
        for k in [k for k in st if st[k] == -1]:
            del self.slot[k]
        self.leaf = {k for k in self.slot if not any(c in self.slot for c in _children(k, dim))}
        marked = [k for k in self.leaf if any(self._nbr_key(k, d) not in self.slot for d in _dirs(dim))]
        nrl, nrr, d2l, d2r = self._ce_sizes()
        if any(p.Bs[a] < max(nrl, nrr) for a in range(dim)):
            raise RuntimeError("adapt_tree: Bs < Nrecon (reconstruction of the neighbours of interface blocks) is not supported")
        leaf_only = all(p.Bs[a] >= d2l and p.Bs[a] >= d2r for a in range(dim))
        if marked:
            if leaf_only:
                self.set_pass_topology(marked)
                sol.coarse_extension_modify(WD, (HVY_BLOCK, 0), True, False)
                sol.waveletReconstruction_CE(WD, (HVY_BLOCK, 0), (HVY_BLOCK, 0))
            else:
                self.set_pass_topology(marked)                                # the lasting interfaces first (adapt_tree.f90:222-228):
                sol.coarse_extension_modify(WD, (HVY_BLOCK, 0), True, False)  # same-level neighbours send coefficients that carry it
                for level in range(self.Jmin, max(k[0] for k in self.slot) + 1):
                    todo = [k for k in marked if k[0] == level]
                    if todo:
                        self.set_pass_topology(todo)
                        sol.waveletReconstruction_CE(WD, (HVY_BLOCK, 0), (HVY_BLOCK, 0))
        # prune_fulltree2leafs + balanceLoad_tree: the leaves move to their slots along the space-filling curve
        keys = sorted(self.leaf)
        new = Forest.from_blocks(dim, self.forest.Jmax, np.array([k[0] for k in keys], dtype=np.int32),
                                 np.array([k[1:] for k in keys], dtype=np.int32), block_dist=self.forest.block_dist, n_ranks=1,
                                 max_blocks=self.forest.max_blocks, periodic=self.forest.periodic)
        hvy, lvl, ixyz, _ = new.active(0)
        src = np.array([self.slot[(int(l), int(x[0]), int(x[1]), int(x[2]))] for l, x in zip(lvl, ixyz)], dtype=np.int32)
        sol.move_blocks(src, hvy.astype(np.int32))
        sol.set_forest(new)
        return new, {"status0": st0, "status": st, "marked": sorted(marked), "leaf_only": leaf_only, "leaf_first": self.leaf_first}
