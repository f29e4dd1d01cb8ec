"""Host-side mask generation for simple geometries (the mask function stays host code in a WABBIT build: CREATE_MASK_meta is called per block
by createMask_tree, LIB/MESH/createMask_tree.f90, and the result is streamed to the device array hvy_mask).

CylinderMask2D: geometry = cylinder / circle of create_mask_2D_ACM (LIB/EQUATION/ACMnew/create_mask.f90:183-320):
  chi    = step_cosine(|x - x_cntr| - R_cyl, h), h = C_smooth * dx_min, dx_min the lattice spacing on Jmax (module_ACM.f90:459-471;
           draw_circle, LIB/EQUATION/insects/module_geometry.f90:315-381; step_cosine4, LIB/HELPER/module_helpers.f90:456-470)
  u_s    = 0, colour = 1 (CREATE_MASK_meta's default, module_physics_metamodule.f90:47-48)
  sponge = p-norm sponge of sponge_2D (LIB/EQUATION/ACMnew/sponge.f90), interior points only
for whole lists of blocks at once (vectorised over blocks)."""
from __future__ import annotations

import numpy as np

from .params import Params


def _step_cosine(x_rel, h):
    out = 0.5 * (1.0 + np.cos((x_rel + h) * np.pi / (2.0 * h)))
    out = np.where(x_rel <= -h, 1.0, out)
    return np.where(x_rel >= h, 0.0, out)


class CylinderMask2D:
    def __init__(self, p: Params, x_cntr=(10.0, 10.0), R_cyl: float = 0.5, C_smooth: float = 1.5, L_sponge: float = 2.0, p_sponge: float = 8.0):
        if p.dim != 2:
            raise ValueError("CylinderMask2D is two-dimensional")
        self.p, self.c, self.R = p, x_cntr, R_cyl
        self.h = C_smooth * min(2.0 ** (-p.Jmax) * p.domain[d] / float(p.Bs[d]) for d in range(2))
        self.L, self.ps = L_sponge, p_sponge

    def _coords(self, level, pos, n_extra):
        p = self.p
        lvl = np.asarray(level, dtype=np.float64)
        out = []
        for d in range(2):
            dx = 2.0 ** (-lvl) * p.domain[d] / float(p.Bs[d])
            x0 = (np.asarray(pos)[:, d] * p.Bs[d]).astype(np.float64) * dx
            out.append(np.arange(p.Bs[d] + n_extra, dtype=np.float64)[None, :] * dx[:, None] + x0[:, None])
        return out

    def chi(self, level, pos, n_extra: int = 0) -> np.ndarray:
        """mask function on the interior points (+ n_extra points behind them): [n, By + n_extra, Bx + n_extra]"""
        x, y = self._coords(level, pos, n_extra)
        dist = np.sqrt((x[:, None, :] - self.c[0]) ** 2 + (y[:, :, None] - self.c[1]) ** 2) - self.R
        return _step_cosine(dist, self.h)

    def fill(self, level, pos) -> np.ndarray:
        """hvy_mask of the listed blocks as createCompleteMaskDirect_tree leaves it: [n, 6, 1, By + 2g, Bx + 2g]"""
        p, g = self.p, self.p.g
        n = len(level)
        m = np.zeros((n, 6, 1, p.Bs[1] + 2 * g, p.Bs[0] + 2 * g))
        m[:, 4] = 1.0
        m[:, 0, 0, g:g + p.Bs[1] + 1, g:g + p.Bs[0] + 1] = self.chi(level, pos, 1)
        if p.use_sponge:
            off = 0.5 * p.domain[0]
            x, y = self._coords(level, pos, 0)
            tmp = -(((x[:, None, :] - off) ** self.ps + (y[:, :, None] - off) ** self.ps) ** (1.0 / self.ps) - off)
            m[:, 5, 0, g:g + p.Bs[1], g:g + p.Bs[0]] = _step_cosine(tmp - 0.5 * self.L, 0.5 * self.L)
        return m

    def keeps(self, level, pos) -> np.ndarray:
        """threshold_mask (coarseningIndicatorMask_tree, LIB/MESH/coarseningIndicator_tree.f90:290-331): True where the mask function is
        not constant over the block's interior"""
        c = self.chi(level, pos).reshape(len(level), -1)
        return ((c > 1.0e-12) & (c < 1.0 - 1.0e-12)).any(axis=1) | ((c.max(axis=1) - c.min(axis=1)) > 1.0e-12)
