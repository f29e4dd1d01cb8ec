"""oracle/adaptive.py -- CPU restatement of WABBIT's adaptive time loop (TEST INFRASTRUCTURE ONLY: imported by tests/ and by
tests/golden/make_golden.py, never by the product).

Reference (paths relative to the reference checkout):
  main time loop                 LIB/MAIN/main.f90:305-443      sync_ghosts_tree -> refine_tree -> timeStep_tree -> adapt_tree -> save
  setInitialCondition_tree       LIB/MESH/setInitialCondition_tree.f90   (read_from_files = 1, adapt_inicond = 1: ONE adapt_tree)
  refine_tree                    LIB/MESH/refine_tree.f90:7-120
  refinementIndicator_tree       LIB/INDICATORS/refinementIndicator_tree.f90:14-257   ("everywhere", "significant")
  respectJmaxJmin_tree           LIB/MESH/respectJmaxJmin_tree.f90
  ensureGradedness_tree          LIB/MESH/ensureGradedness_tree.f90     (refinement part: a block whose finer neighbour refines, refines)
  refinement_execute_tree        LIB/MESH/refinementExecute.f90:1-120   (refineBlock: prediction of the ghosted block)
  componentWiseNorm_tree         LIB/OPERATORS/componentWiseNorm_tree.f90 (Linfty over the interiors of the leaves)

Everything numerical is delegated to the restatements in oracle.py / fulltree.py / orc_*.c.

PINNED by the reference's own regression fixtures TESTING/acm/3vortices/3vorticesAdaptFD4_CDF4{0,2} (tests/test_oracle_adaptive.py):
the stored grid after adapt_inicond (t = 10) and after the 2281 adaptive time steps to t = 15 -- block lists, refinement statuses
and iteration counter identical, fields to round-off.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

import oracle as O
import fulltree as FT

REF_STAY = FT.REF_STAY


def step_cosine(x_rel: np.ndarray, h: float) -> np.ndarray:
    """step_cosine3 / step_cosine4 (LIB/HELPER/module_helpers.f90:436-470) of x_rel = x - t"""
    out = 0.5 * (1.0 + np.cos((x_rel + h) * np.pi / (2.0 * h)))
    out = np.where(x_rel <= -h, 1.0, out)
    return np.where(x_rel >= h, 0.0, out)


class CylinderMask2D:
    """The six mask components of create_mask_2D_ACM (LIB/EQUATION/ACMnew/create_mask.f90:183-320) for geometry = cylinder (a circle,
    draw_circle, LIB/EQUATION/insects/module_geometry.f90:315-381, cosine smoothing of width C_smooth * dx_min with dx_min the lattice
    spacing on Jmax, module_ACM.f90:459-471) plus the p-norm sponge (sponge_2D, sponge.f90), as CREATE_MASK_meta ("all-parts") leaves them
    on a block: [chi, 0, 0, 0, colour = 1, sponge] at the interior points (the sponge) / the interior and the first upper ghost point (chi);
    zero elsewhere.  A pure function of the block's position: createMask_tree in 2-D always draws all parts directly."""

    def __init__(self, p: O.Params, center=(10.0, 10.0), radius: float = 0.5, C_smooth: float = 1.5, use_sponge: bool = True, L_sponge: float = 2.0,
                 p_sponge: float = 8.0):
        self.p, self.center, self.radius = p, center, radius
        dx_min = min(2.0 ** (-p.Jmax) * p.domain[d] / float(p.Bs[d]) for d in range(p.dim))
        self.h = dx_min * C_smooth
        self.use_sponge, self.L, self.ps = use_sponge, L_sponge, p_sponge

    def block(self, level: int, ixyz) -> np.ndarray:
        p, g = self.p, self.p.g
        Bx, By = p.Bs[0], p.Bs[1]
        m = np.zeros((6, 1, By + 2 * g, Bx + 2 * g))
        m[4] = 1.0
        dx = [2.0 ** (-level) * p.domain[d] / float(p.Bs[d]) for d in range(2)]
        x0 = [float(int(ixyz[d]) * p.Bs[d]) * dx[d] for d in range(2)]
        x = np.arange(0, Bx + 1, dtype=np.float64) * dx[0] + x0[0]
        y = np.arange(0, By + 1, dtype=np.float64) * dx[1] + x0[1]
        dist = np.sqrt((x[None, :] - self.center[0]) ** 2 + (y[:, None] - self.center[1]) ** 2) - self.radius
        m[0, 0, g:g + By + 1, g:g + Bx + 1] = step_cosine(dist, self.h)
        if self.use_sponge:
            off = 0.5 * p.domain[0]
            xs, ys = x[:Bx] - off, y[:By] - off
            tmp = -((xs[None, :] ** self.ps + ys[:, None] ** self.ps) ** (1.0 / self.ps) - off)
            m[5, 0, g:g + By, g:g + Bx] = step_cosine(tmp - 0.5 * self.L, 0.5 * self.L)
        return m

    def keeps(self, level: int, ixyz) -> bool:
        """coarseningIndicatorMask_tree: the mask function varies over the block's interior"""
        p, g = self.p, self.p.g
        chi = self.block(level, ixyz)[0, 0, g:g + p.Bs[1], g:g + p.Bs[0]]
        return bool(((chi > 1.0e-12) & (chi < 1.0 - 1.0e-12)).any() or (chi.max() - chi.min()) > 1.0e-12)


class SphereMask3D:
    """A sphere drawn by draw_sphere (LIB/EQUATION/insects/module_geometry.f90, 'sphere-fixed' of create_mask_3D_ACM, create_mask.f90:6-174)
    with cosine smoothing of width C_smooth * dx_min, optionally translating with a constant velocity (SURVEY 8d config 4's synthetic
    stand-in for a moving body: centre(t) = centre0 + velocity * t, solid velocity u_s = velocity).  Components [chi, u_s(3), colour = 1,
    sponge = 0]; chi on the interior points and the first upper ghost point, like draw_sphere's loop bounds."""

    def __init__(self, p: O.Params, center=(0.5, 0.5, 0.5), radius: float = 0.15, velocity=(0.0, 0.0, 0.0), C_smooth: float = 1.5):
        self.p, self.c0, self.R, self.v = p, np.asarray(center, dtype=np.float64), radius, np.asarray(velocity, dtype=np.float64)
        self.h = C_smooth * min(2.0 ** (-p.Jmax) * p.domain[d] / float(p.Bs[d]) for d in range(3))

    def chi(self, level: int, ixyz, time: float, n_extra: int = 0) -> np.ndarray:
        p = self.p
        c = self.c0 + self.v * time
        ax = []
        for d in range(3):
            dx = 2.0 ** (-level) * p.domain[d] / float(p.Bs[d])
            x0 = float(int(ixyz[d]) * p.Bs[d]) * dx
            ax.append(np.arange(0, p.Bs[d] + n_extra, dtype=np.float64) * dx + x0)
        dist = np.sqrt((ax[0][None, None, :] - c[0]) ** 2 + (ax[1][None, :, None] - c[1]) ** 2 + (ax[2][:, None, None] - c[2]) ** 2) - self.R
        return step_cosine(dist, self.h)

    def block(self, level: int, ixyz, time: float = 0.0) -> np.ndarray:
        p, g = self.p, self.p.g
        B = p.Bs
        m = np.zeros((6, B[2] + 2 * g, B[1] + 2 * g, B[0] + 2 * g))
        m[4] = 1.0
        for a in range(3):
            m[1 + a] = self.v[a]
        m[0, g:g + B[2] + 1, g:g + B[1] + 1, g:g + B[0] + 1] = self.chi(level, ixyz, time, 1)
        return m

    def keeps(self, level: int, ixyz, time: float = 0.0) -> bool:
        chi = self.chi(level, ixyz, time)
        return bool(((chi > 1.0e-12) & (chi < 1.0 - 1.0e-12)).any() or (chi.max() - chi.min()) > 1.0e-12)


class AdaptiveRun:
    """State of one adaptive simulation: leaf grid, ghosted data [nb, nc, nz, ny, nx], refinement status per leaf, time, iteration."""

    def __init__(self, p: O.Params, wavelet: str, grid: O.Grid, u: np.ndarray, time: float, iteration: int, eps: float, Jmin: int = 1,
                 refinement_indicator: str = "everywhere", use_coarse_extension: Optional[bool] = None,
                 use_security_zone: Optional[bool] = None, fd_half_width: int = 2, force_maxlevel_dealiasing: bool = False,
                 thresh_comp=None, eps_normalized: bool = True, eps_norm: str = "Linfty", mask=None, threshold_mask: bool = False,
                 mask_time_dependent: bool = False):
        self.p, self.w, self.grid, self.u = p, O.setup_wavelet(wavelet), grid, u
        self.time, self.iteration, self.eps, self.Jmin = time, iteration, eps, Jmin
        self.refinement_indicator = refinement_indicator
        self.use_ce = bool(self.w.lifted) if use_coarse_extension is None else use_coarse_extension
        self.use_sz = bool(self.w.lifted) if use_security_zone is None else use_security_zone
        self.fd_half_width = fd_half_width
        self.dealias = force_maxlevel_dealiasing
        self.thresh_comp = thresh_comp
        self.eps_normalized, self.eps_norm = eps_normalized, eps_norm
        self.mask, self.threshold_mask = mask, threshold_mask        # mask: object with block(level, ixyz[, t]) and keeps(level, ixyz[, t])
        self.mask_time_dependent = mask_time_dependent
        self.status = np.zeros(grid.n, dtype=np.int64)
        self.adapted_once = False
        self.log = []

    def _mt(self):
        return (self.time,) if self.mask_time_dependent else ()

    # ------------------------------------------------------------------------------------------------------------------
    def _nbr(self):
        """hvy_neighbor of the current grid (updateMetadata_tree), computed once per grid"""
        if getattr(self, "_nbr_of", None) is not self.grid:
            self._nbr_tab, self._nbr_of = O.neighbor_table168(self.grid, self.p.Jmax), self.grid
        return self._nbr_tab

    def sync_ghosts_tree(self):
        """sync_ghosts_tree: all g ghost nodes, restriction through the HD filter for lifted wavelets (idempotent: repeated calls on
        unchanged data are skipped)"""
        if getattr(self, "_synced", None) is self.u:
            return
        O.sync_ghosts_leaf(self.grid, self.p, self.u, self._nbr(), self.p.g, self.p.g, self.w.X, bool(self.w.lifted),
                           ignore_filter=not self.w.lifted, w=self.w)
        self._synced = self.u

    def norm(self) -> np.ndarray:
        """componentWiseNorm_tree(..., "Linfty") on the leaves' interiors; values <= 1e-9 become 1 (coarseningIndicator_tree.f90:66-68)"""
        if not self.eps_normalized:
            return np.ones(self.u.shape[1])
        assert self.eps_norm == "Linfty"
        it = (slice(None), slice(None)) + O.interior(self.p)
        n = np.abs(self.u[it]).max(axis=(0, 2, 3, 4))
        if self.thresh_comp is not None:                                 # componentWiseNorm_tree.f90:119-163: 0 not computed, >= 2 joint
            tc = np.asarray(self.thresh_comp)
            n = np.where(tc == 0, -1.0, n)
            for l in range(2, int(tc.max()) + 1):
                if (tc == l).any():
                    n[tc == l] = n[tc == l].max()
        n[n <= 1.0e-9] = 1.0
        return n

    # ------------------------------------------------------------------------------------------------------------------
    def adapt_tree(self):
        g, self.u, info = FT.adapt_tree(self.p, self.w, self.grid, self.u, self.eps, Jmin=self.Jmin, norm=self.norm(), eps_norm=self.eps_norm,
                                        thresh_comp=self.thresh_comp, level_ref=self.p.Jmax, force_maxlevel_dealiasing=self.dealias,
                                        fd_half_width=self.fd_half_width, use_security_zone=self.use_sz, use_coarse_extension=self.use_ce,
                                        mask_keeps=(lambda k: self.mask.keeps(k[0], k[1:], *self._mt())) if (self.mask is not None and self.threshold_mask) else None)
        self.grid = g
        st = info["status"]
        self.status = np.array([st[(int(l),) + tuple(int(v) for v in x)] for l, x in zip(g.level, g.ixyz)], dtype=np.int64)
        self.adapted_once = True
        self.sync_ghosts_tree()                                          # adapt_tree.f90:258
        return info

    # ------------------------------------------------------------------------------------------------------------------
    def refine_flags(self, indicator: str) -> np.ndarray:
        g, dim = self.grid, self.grid.dim
        if indicator == "everywhere":
            flag = np.ones(g.n, dtype=np.int64)
        elif indicator == "significant":
            assert not np.isin(self.status, (-1, 1)).any()               # abort(241119)
            flag = np.where(self.status == 0, 1, 0).astype(np.int64)     # 0 -> +1, REF_UNSIGNIFICANT_STAY -> 0
        else:
            raise ValueError(indicator)
        flag[(flag == 1) & (g.level >= self.p.Jmax)] = 0                 # respectJmaxJmin_tree
        if indicator != "everywhere":                                    # ensureGradedness_tree: monotone 0 -> +1, order-independent
            key = {(int(l),) + tuple(int(v) for v in x): b for b, (l, x) in enumerate(zip(g.level, g.ixyz))}
            changed = True
            while changed:
                changed = False
                for k, b in key.items():
                    if flag[b] != 0:
                        continue
                    for d in FT.dirs(dim):
                        nk = FT.nbr_key(k, d, dim)
                        if nk in key:
                            continue                                     # same-level neighbour
                        hit = False
                        for c in FT.children(nk, dim):                   # finer neighbours touching k across d
                            if c in key and flag[key[c]] == 1 and all(
                                    (d[a] == 0) or (d[a] > 0 and (c[1 + a] & 1) == 0) or (d[a] < 0 and (c[1 + a] & 1) == 1) for a in range(dim)):
                                hit = True
                        if hit:
                            flag[b] = 1
                            changed = True
                            break
        return flag

    def refine_tree(self, indicator: Optional[str] = None):
        """refine_tree; the caller has synchronised all ghost nodes (main.f90:314)"""
        indicator = self.refinement_indicator if indicator is None else indicator
        if indicator == "significant" and not self.adapted_once:
            indicator = "everywhere"                                     # main.f90:322
        flag = self.refine_flags(indicator)
        g, p, dim = self.grid, self.p, self.grid.dim
        lev, ixyz, data = [], [], []
        for b in range(g.n):
            if flag[b] != 1:
                lev.append(int(g.level[b]))
                ixyz.append(tuple(int(v) for v in g.ixyz[b]))
                data.append(self.u[b])
                continue
            d = O.refine_block(self.w.X, p, self.u[b])                   # refineBlock: daughters in treecode digit order
            k = (int(g.level[b]),) + tuple(int(v) for v in g.ixyz[b])
            for c, ck in enumerate(FT.children(k, dim)):
                lev.append(ck[0])
                ixyz.append(ck[1:])
                data.append(d[self._daughter_slot(c, dim)])
        order = sorted(range(len(lev)), key=lambda i: (lev[i],) + tuple(ixyz[i]))
        self.grid = O.Grid(level=np.array([lev[i] for i in order], dtype=np.int64), ixyz=np.array([ixyz[i] for i in order], dtype=np.int64), dim=dim)
        self.u = np.ascontiguousarray(np.stack([data[i] for i in order]))
        self.status = np.zeros(self.grid.n, dtype=np.int64)
        return int(flag.sum())

    @staticmethod
    def _daughter_slot(c: int, dim: int) -> int:
        """FT.children numbers daughters by qx + 2 qy + 4 qz; orc_refine_block stores them by treecode digit, qy + 2 qx + 4 qz"""
        return ((c >> 1) & 1) + 2 * (c & 1) + 4 * ((c >> 2) & 1)

    # ------------------------------------------------------------------------------------------------------------------
    def time_step(self):
        p, g = self.p, self.grid
        nbr = self._nbr()
        self._synced = None                                              # the step changes the data

        def sync(h):
            O.sync_ghosts_leaf(g, p, h, nbr, p.g_rhs, p.g_rhs, self.w.X, bool(self.w.lifted), ignore_filter=True)
        work = np.zeros((p.butcher.shape[0] + 1,) + self.u.shape)
        mask, mask_at = None, None
        if self.mask is not None and self.mask_time_dependent:
            mask_at = lambda t: np.stack([self.mask.block(int(l), x, t) for l, x in zip(g.level, g.ixyz)])
        elif self.mask is not None:
            mask = np.stack([self.mask.block(int(l), x) for l, x in zip(g.level, g.ixyz)])                             # createMask_tree
        dt = O.rk_generic(g, p, self.u, work, self.time, mask=mask, sync=sync, mask_at=mask_at)
        self.time += dt
        self.iteration += 1
        return dt

    def adaptive_inicond(self, inicond):
        """setInitialCondition_tree with inicond_grid_from_file = "no" and adapt_inicond = 1 (setInitialCondition_tree.f90:97-130): the caller
        starts from the equidistant grid on Jini with the initial condition set; then, until the number of blocks stops changing (at most
        Jmax - Jmin times): refine everywhere, set the initial condition on the new grid, adapt_tree.  inicond(run) fills run.u."""
        n_old, it = 9999999, 0
        while self.grid.n != n_old and it < self.p.Jmax - self.Jmin:
            n_old = self.grid.n
            self.refine_tree("everywhere")
            inicond(self)
            self.adapt_tree()
            it += 1
        return it

    def step(self):
        """one pass of the main loop (main.f90:305-425)"""
        self.sync_ghosts_tree()
        self.refine_tree()
        nb_rhs = self.grid.n
        dt = self.time_step()
        self.adapt_tree()
        self.log.append((self.iteration, self.time, nb_rhs, self.grid.n, dt))
        return dt
