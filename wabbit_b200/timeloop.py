"""The adaptive time loop around the device path: what LIB/MAIN/main.f90:305-443 does per iteration with `adapt_tree = 1`,

    sync_ghosts_tree -> refine_tree(refinement_indicator) -> timeStep_tree (RungeKuttaGeneric) -> adapt_tree

with the heavy data resident on the GPU (WabbitGPU) and the light data (grid, refinement flags) on the host.  The ghost nodes are never
stored on the device: every consumer (refineBlock, the stage kernels, the wavelet kernels) resolves them on the fly, so the two
sync_ghosts_tree calls of the reference loop have no counterpart here.

Host light-data logic of this module (stays in host Fortran in a WABBIT build; restated here for the drivers and tests):
  refinementIndicator_tree    LIB/INDICATORS/refinementIndicator_tree.f90:14-257   "everywhere", "significant"
  respectJmaxJmin_tree        LIB/MESH/respectJmaxJmin_tree.f90
  ensureGradedness_tree       LIB/MESH/ensureGradedness_tree.f90 (refinement half: a block whose finer neighbour refines, refines too)
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .forest import Forest
from .solver import HVY_MASK, WabbitGPU

REF_UNSIGNIFICANT_STAY = 9      # module_globals.f90:28


def refinement_flags(forest: Forest, indicator: str, status: Optional[np.ndarray], Jmax: int) -> np.ndarray:
    """+1 / 0 per block in the global space-filling-curve order (forest.active(0) on one rank): refinementIndicator_tree +
    respectJmaxJmin_tree + ensureGradedness_tree.
    `status`: lgt_block(:, IDX_REFINE_STS) left by the last adapt_tree (0 significant, 9 REF_UNSIGNIFICANT_STAY)."""
    parts = [forest.active(r) for r in range(forest.n_ranks)]          # global order = rank-major order of the active lists
    lvl = np.concatenate([q[1] for q in parts]).astype(np.int64)
    pos = np.concatenate([q[2] for q in parts]).astype(np.int64)
    n, dim = len(lvl), forest.dim
    if indicator == "everywhere":
        flag = np.ones(n, np.int32)
    elif indicator == "significant":
        if status is None:
            raise ValueError('refinement indicator "significant" needs the refinement status of the last adapt_tree')
        if np.isin(status, (-1, 1)).any():
            raise RuntimeError("241119: I am very confused by what is going on here and do not like it!")
        flag = (np.asarray(status) == 0).astype(np.int32)
    else:
        raise ValueError(f"refinement indicator {indicator!r} is not supported")
    flag[lvl >= Jmax] = 0
    if indicator == "everywhere":
        return flag
    # gradedness: a block (level L, flag 0) refines if a finer neighbour (level L+1) refines.  Propagate upwards from the finest level:
    # every refining block marks the (up to 3^dim - 1) leaves of level L-1 that touch it; vectorised over the refining blocks, lookups
    # by position code (sorted array + searchsorted), repeated until nothing changes
    from .fulltree import _pack
    code = _pack(lvl, pos)
    order = np.argsort(code)
    code_sorted = code[order]
    dirs = [(dx, dy, dz) for dz in ((-1, 0, 1) if dim == 3 else (0,)) for dy in (-1, 0, 1) for dx in (-1, 0, 1) if (dx, dy, dz) != (0, 0, 0)]
    front = np.flatnonzero((flag == 1) & (lvl > 0))
    while len(front):
        L = lvl[front]
        mask = (np.int64(1) << L) - 1
        hit = []
        for d in dirs:
            q = (pos[front] + np.array(d, dtype=np.int64)[None, :]) & mask[:, None]
            if dim == 2:
                q[:, 2] = 0
            ck = _pack(L - 1, q >> 1)
            k = np.minimum(np.searchsorted(code_sorted, ck), n - 1)
            ok = code_sorted[k] == ck
            hit.append(order[k[ok]])
        cand = np.unique(np.concatenate(hit))
        new = cand[flag[cand] == 0]
        new = new[lvl[new] < Jmax]                       # (always true: a leaf with a finer neighbour is below Jmax)
        flag[new] = 1
        front = new[lvl[new] > 0]
    return flag


class AdaptiveLoop:
    """One simulation with adapt_tree = 1 on one GPU."""

    def __init__(self, sol: WabbitGPU, forest: Forest, time: float = 0.0, iteration: int = 0, refinement_indicator: Optional[str] = None,
                 thresh_comp=None, mask=None, threshold_mask: bool = False):
        """mask: host geometry object with fill(level, pos) -> hvy_mask rows and keeps(level, pos) -> bool (wabbit_b200.mask)"""
        self.sol, self.forest, self.time, self.iteration = sol, forest, time, iteration
        self.mask, self.threshold_mask = mask, threshold_mask
        self.mask_time_dependent = mask is not None and bool(np.any(getattr(mask, "v", 0.0)))
        if mask is not None and getattr(mask, "analytic", False):
            mask.attach(sol)                                 # evaluated inside the stage kernel: no hvy_mask traffic
        p = sol.params
        self.indicator = p.refinement_indicator if refinement_indicator is None else refinement_indicator
        self.thresh_comp = thresh_comp
        self.status: Optional[np.ndarray] = None
        self.log = []

    def adapt_tree(self):
        p = self.sol.params
        self.forest, n0, n1 = self.sol.adapt_tree(self.forest, eps=p.eps, eps_normalized=p.eps_normalized, eps_norm=p.eps_norm, Jmin=p.Jmin,
                                                  force_maxlevel_dealiasing=p.force_maxlevel_dealiasing, thresh_comp=self.thresh_comp,
                                                  mask_keeps=self._mask_keeps() if (self.mask is not None and self.threshold_mask) else None,
                                                  full_tree=True)     # the reference's algorithm for every wavelet (adapt_tree.f90:11-260)
        self.status = self.sol.refinement_status
        return n0, n1

    def _mask_keeps(self):
        if self.mask_time_dependent:
            return lambda level, pos: self.mask.keeps(level, pos, self.time)
        return self.mask.keeps

    def refine_tree(self):
        ind = self.indicator
        if ind == "significant" and self.status is None:
            ind = "everywhere"                                   # main.f90:322: adapt_tree was not called yet
        flags = None if ind == "everywhere" else refinement_flags(self.forest, ind, self.status, self.sol.params.Jmax)
        self.forest = self.sol.refine_tree(self.forest, flags)
        self.status = None
        return self.forest.n_blocks

    def createMask_tree(self):
        """createMask_tree on the current grid (2-D: always all parts, drawn directly): host geometry -> hvy_mask on the device"""
        if self.mask is None or getattr(self.mask, "analytic", False):
            return
        hvy, lvl, pos, _ = self.forest.active(0)
        host = np.zeros((int(hvy.max()),) + self.sol.host_shape(self.sol.params.n_mask)[1:])     # rows 1 .. max hvy id only
        host[hvy - 1] = self.mask.fill(lvl, pos)
        self.sol.upload(host, HVY_MASK, 0, hvy_ids=hvy)

    def adaptive_inicond(self, set_inicond) -> int:
        """setInitialCondition_tree (LIB/MESH/setInitialCondition_tree.f90:97-130) with adapt_inicond = 1 on an auto-generated grid: until the
        number of blocks stops changing (at most Jmax - Jmin times) refine everywhere, set the initial condition, adapt_tree.
        set_inicond(loop) uploads the initial condition for loop.forest."""
        p = self.sol.params
        n_old, it = 9999999, 0
        while self.forest.n_blocks != n_old and it < p.Jmax - p.Jmin:
            n_old = self.forest.n_blocks
            self.forest = self.sol.refine_tree(self.forest, None)
            set_inicond(self)
            self.adapt_tree()
            it += 1
        return it

    stats_dir: Optional[str] = None      # where the *.t files go; None: no statistics

    def statistics(self, dt: float) -> Optional[dict]:
        """main.f90:388-397: statistics_wrapper every nsave_stats iterations / tsave_stats time units, after the time step and before the grid
        is coarsened; the rows go to the *.t files of `stats_dir` (wabbit_b200.tfiles)"""
        from . import tfiles
        p = self.sol.params
        if self.stats_dir is None or not tfiles.statistics_due(self.iteration, self.time, p.nsave_stats, p.tsave_stats):
            return None
        if self.mask is not None and getattr(self.mask, "analytic", False):
            self.mask.fill_device(self.sol, self.time)       # the statistics kernel reads hvy_mask; the stage kernel does not need it
        else:
            self.createMask_tree()
        stats = self.sol.statistics_ACM(self.time, with_divergence=True, with_vorticity=True)
        _, lvl, _, _ = self.forest.active(0)
        dx_min = min(2.0 ** (-int(np.max(lvl))) * p.domain[d] / float(p.Bs[d]) for d in range(p.dim))
        tfiles.write_statistics_acm(stats, self.time, dt, p, dx_min, self.stats_dir)
        return stats

    def step(self) -> float:
        nb_rhs = self.refine_tree()
        self.createMask_tree()
        self.time, self.iteration, dt = self.sol.timeStep_tree(self.time, self.iteration)
        self.statistics(dt)
        self.adapt_tree()
        self.log.append((self.iteration, self.time, nb_rhs, self.forest.n_blocks, dt))
        return dt


class DistributedAdaptiveLoop:
    """The same loop with the blocks partitioned over ranks (wabbit_b200.multi.DistributedWabbit: halo blocks + block transport); the light
    data -- flags, statuses, the mask indicator -- are replicated, so every rank takes the same grid decisions."""

    def __init__(self, drv, time: float = 0.0, iteration: int = 0, refinement_indicator: Optional[str] = None, thresh_comp=None, mask=None,
                 threshold_mask: bool = False):
        self.drv, self.time, self.iteration = drv, time, iteration
        p = drv.sol.params
        self.indicator = p.refinement_indicator if refinement_indicator is None else refinement_indicator
        self.thresh_comp, self.mask, self.threshold_mask = thresh_comp, mask, threshold_mask
        self.mask_time_dependent = mask is not None and bool(np.any(getattr(mask, "v", 0.0)))
        if mask is not None:
            if not getattr(mask, "analytic", False):
                raise ValueError("DistributedAdaptiveLoop: only masks evaluated on the device (wabbit_b200.mask.SphereMask3D) are supported")
            mask.attach(drv.sol)
        self.status: Optional[np.ndarray] = None
        self.log = []

    def adapt_tree(self):
        p = self.drv.sol.params
        keeps = None
        if self.mask is not None and self.threshold_mask:
            keeps = (lambda level, pos: self.mask.keeps(level, pos, self.time)) if self.mask_time_dependent else self.mask.keeps
        _, n0, n1 = self.drv.adapt_tree(eps=p.eps, eps_normalized=p.eps_normalized, eps_norm=p.eps_norm, Jmin=p.Jmin,
                                        force_maxlevel_dealiasing=p.force_maxlevel_dealiasing, thresh_comp=self.thresh_comp, mask_keeps=keeps,
                                        full_tree=True)
        self.status = self.drv.refinement_status
        return n0, n1

    def step(self) -> float:
        ind = self.indicator
        if ind == "significant" and self.status is None:
            ind = "everywhere"
        flags = None if ind == "everywhere" else refinement_flags(self.drv.forest, ind, self.status, self.drv.sol.params.Jmax)
        nb_rhs = self.drv.refine_tree(flags).n_blocks
        self.status = None
        self.time, self.iteration, dt = self.drv.timeStep_tree(self.time, self.iteration)
        self.adapt_tree()
        self.log.append((self.iteration, self.time, nb_rhs, self.drv.forest.n_blocks, dt))
        return dt
