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
        if p.penalization:                 # create_mask_2D_ACM draws the geometry only with penalization = 1; the sponge is independent of it
            m[:, 0, 0, g:g + p.Bs[1] + 1, g:g + p.Bs[0] + 1] = self.chi(level, pos, 1)
        if p.use_sponge:
            off = 0.5 * p.domain[0]
            x, y = self._coords(level, pos, 0)
            tmp = -(((x[:, None, :] - off) ** self.ps + (y[:, :, None] - off) ** self.ps) ** (1.0 / self.ps) - off)
            m[:, 5, 0, g:g + p.Bs[1], g:g + p.Bs[0]] = _step_cosine(tmp - 0.5 * self.L, 0.5 * self.L)
        return m

    def fill_device(self, sol, time: float = 0.0):
        """the same six components evaluated on the device straight into the resident hvy_mask (wgpu_create_mask): no host array, no upload"""
        sol.create_mask_device(time, "cylinder", self.c, (0.0, 0.0), self.R, self.h, self.L, self.ps)

    def keeps(self, level, pos) -> np.ndarray:
        """threshold_mask (coarseningIndicatorMask_tree, LIB/MESH/coarseningIndicator_tree.f90:290-331): True where the mask function is
        not constant over the block's interior"""
        if not self.p.penalization:
            return np.zeros(len(level), bool)          # chi is identically zero
        c = self.chi(level, pos).reshape(len(level), -1)
        return ((c > 1.0e-12) & (c < 1.0 - 1.0e-12)).any(axis=1) | ((c.max(axis=1) - c.min(axis=1)) > 1.0e-12)


class SphereMask3D:
    """A sphere (draw_sphere, 'sphere-fixed' of create_mask_3D_ACM) with cosine smoothing, optionally translating with a constant velocity
    (SURVEY 8d config 4's synthetic moving body: centre(t) = centre0 + velocity * t, u_s = velocity).  On the device the mask is evaluated
    inside the stage kernel (`attach`: wgpu_set_mask_sphere) -- nothing is generated, uploaded or read per stage; the host only needs
    `keeps` for threshold_mask (light data)."""
    analytic = True

    def __init__(self, p: Params, center=(0.5, 0.5, 0.5), radius: float = 0.15, velocity=(0.0, 0.0, 0.0), C_smooth: float = 1.5):
        if p.dim != 3:
            raise ValueError("SphereMask3D is three-dimensional")
        self.p, self.c0, self.R, self.v = p, np.asarray(center, dtype=np.float64), radius, np.asarray(velocity, dtype=np.float64)
        self.h = C_smooth * min(2.0 ** (-p.Jmax) * p.domain[d] / float(p.Bs[d]) for d in range(3))

    def attach(self, sol):
        sol.set_mask_sphere(self.c0, self.v, self.R, self.h)

    def fill_device(self, sol, time: float = 0.0):
        """the six mask components at `time` written into the resident hvy_mask on the device (wgpu_create_mask) -- for consumers that read
        hvy_mask (statistics, a stage kernel without the in-kernel mask)"""
        sol.create_mask_device(time, "sphere", self.c0, self.v, self.R, self.h)

    def keeps(self, level, pos, time: float = 0.0) -> np.ndarray:
        """threshold_mask: True where the mask function varies over the block's interior.  Blocks whose bounding box does not come within
        the smoothing width of the sphere's surface are constant and are skipped; the others are evaluated point by point."""
        p = self.p
        level, pos = np.asarray(level), np.asarray(pos)
        c = self.c0 + self.v * time
        n = len(level)
        dx = np.stack([2.0 ** (-level.astype(np.float64)) * p.domain[d] / float(p.Bs[d]) for d in range(3)], axis=1)
        lo = pos[:, :3] * np.asarray(p.Bs[:3])[None, :] * dx
        hi = lo + (np.asarray(p.Bs[:3])[None, :] - 1) * dx
        near = np.clip(c[None, :], lo, hi)                                        # closest / farthest lattice-box points to the centre
        far = np.where(np.abs(lo - c[None, :]) > np.abs(hi - c[None, :]), lo, hi)
        dmin = np.sqrt(((near - c[None, :]) ** 2).sum(axis=1)) - self.R
        dmax = np.sqrt(((far - c[None, :]) ** 2).sum(axis=1)) - self.R
        cand = np.flatnonzero((dmin < self.h + dx.max(axis=1)) & (dmax > -self.h - dx.max(axis=1)))
        out = np.zeros(n, bool)
        for s0 in range(0, len(cand), 128):                          # point-by-point evaluation, 128 candidate blocks at a time
            ii = cand[s0:s0 + 128]
            ax = [np.arange(p.Bs[d], dtype=np.float64)[None, :] * dx[ii, d][:, None]
                  + ((pos[ii, d] * p.Bs[d]).astype(np.float64) * dx[ii, d])[:, None] for d in range(3)]
            dist = np.sqrt((ax[0][:, None, None, :] - c[0]) ** 2 + (ax[1][:, None, :, None] - c[1]) ** 2 + (ax[2][:, :, None, None] - c[2]) ** 2) - self.R
            chi = _step_cosine(dist, self.h).reshape(len(ii), -1)
            out[ii] = ((chi > 1.0e-12) & (chi < 1.0 - 1.0e-12)).any(axis=1) | ((chi.max(axis=1) - chi.min(axis=1)) > 1.0e-12)
        return out


def mask_from_ini(path: str, p: Params):
    """The mask generator for the [VPM] / [Sponge] sections of a WABBIT parameter file (READ_PARAMETERS_ACM, LIB/EQUATION/ACMnew/
    module_ACM.f90:290-345: geometry, x_cntr, R_cyl = 0.5, C_smooth = 1.5, smoothing_type = cos; sponge_type, L_sponge, p_sponge = 20) or None
    with neither penalization nor a sponge.  The sponge does not depend on penalization (create_mask.f90: `if (params_acm%use_sponge)`), so
    penalization = 0 with use_sponge = 1 returns a generator whose chi is zero.  Supported: geometry = cylinder / circle (2-D, cosine
    smoothing, p-norm sponge) and sphere-fixed (3-D, no sponge)."""
    from .params import IniFile
    if not (p.penalization or p.use_sponge):
        return None
    ini = IniFile(path)
    if not p.penalization:
        if p.dim != 2 or ini.string("Sponge", "sponge_type", "rect").strip().lower() != "p-norm":
            raise ValueError("use_sponge = 1 without penalization: only the 2-D p-norm sponge is supported")
        return CylinderMask2D(p, L_sponge=ini.real("Sponge", "L_sponge", 0.0), p_sponge=ini.real("Sponge", "p_sponge", 20.0))
    geometry = ini.string("VPM", "geometry", "cylinder").strip().lower()
    x_cntr = ini.vector("VPM", "x_cntr", [0.5 * p.domain[0], 0.5 * p.domain[1], 0.5 * p.domain[2]])
    x_cntr = (list(x_cntr) + [0.0, 0.0, 0.0])[:3]
    R = ini.real("VPM", "R_cyl", 0.5)
    C_smooth = ini.real("VPM", "C_smooth", 1.5)
    smoothing = ini.string("VPM", "smoothing_type", "cos").strip().lower()
    if smoothing not in ("cos", "cosine"):
        raise ValueError(f"smoothing_type {smoothing!r} is not supported (cosine only)")
    if geometry in ("cylinder", "circle"):
        if p.use_sponge and ini.string("Sponge", "sponge_type", "rect").strip().lower() != "p-norm":
            raise ValueError("only the p-norm sponge is supported")
        return CylinderMask2D(p, x_cntr=tuple(x_cntr[:2]), R_cyl=R, C_smooth=C_smooth, L_sponge=ini.real("Sponge", "L_sponge", 0.0),
                              p_sponge=ini.real("Sponge", "p_sponge", 20.0))
    if geometry == "sphere-fixed":
        return SphereMask3D(p, center=tuple(x_cntr), radius=R, C_smooth=C_smooth)
    raise ValueError(f"geometry {geometry!r} is not supported")
