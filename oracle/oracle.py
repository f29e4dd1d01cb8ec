"""CPU oracle for the WABBIT block hot path (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
``wabbit_b200`` never does; it fails loudly when its CUDA library is missing.

The arithmetic lives in ``orc_*.c`` (plain C, ``-ffp-contract=off``), restating the
reference Fortran statement by statement; this file holds the tree-level loops
(block loops, ghost synchronisation on the grid, Runge-Kutta driver, time-step
control) in NumPy, citing the reference file:line each part follows.

Array convention: ``hvy[b, c, iz, iy, ix]`` C-ordered == Fortran ``hvy(ix,iy,iz,c,b)``.
2-D runs use ``nz = 1``.
"""
from __future__ import annotations

import ctypes as C
import functools
import math
import os
import subprocess
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRCS = ["orc_acm.c", "orc_tree.c", "orc_wavelet.c", "orc_sync.c"]


def _lib_path(fast: bool) -> str:
    return os.path.join(_HERE, "liboracle_fast.so" if fast else "liboracle.so")


def build(force: bool = False) -> None:
    """Compile the C restatement: parity build (no FMA contraction) and timing build."""
    srcs = [os.path.join(_HERE, s) for s in _SRCS if os.path.exists(os.path.join(_HERE, s))]
    for fast in (False, True):
        out = _lib_path(fast)
        if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in srcs):
            continue
        flags = ["-O3", "-march=native", "-fopenmp"] if fast else ["-O2", "-ffp-contract=off", "-fopenmp"]
        subprocess.check_call(["gcc", "-shared", "-fPIC", "-std=c11", *flags, "-o", out, *srcs, "-lm"])


class AcmParams(C.Structure):
    _fields_ = [("dim", C.c_int32), ("fd", C.c_int32), ("skew", C.c_int32), ("penalization", C.c_int32),
                ("use_sponge", C.c_int32), ("pad_", C.c_int32),
                ("c0", C.c_double), ("nu", C.c_double), ("gamma_p", C.c_double), ("C_eta", C.c_double),
                ("C_sponge", C.c_double), ("u_mean_set", C.c_double * 3),
                ("CFL", C.c_double), ("CFL_eta", C.c_double), ("CFL_nu", C.c_double)]


FD_IDS = {"FD_2nd_central": 2, "FD_4th_central": 4, "FD_6th_central": 6, "FD_4th_central_optimized": 40}

_libs: Dict[bool, C.CDLL] = {}
_dp = C.c_void_p                      # double* arguments: takes the typed pointers of _p and the plain addresses of _pb
_dpp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def lib(fast: bool = False) -> C.CDLL:
    if fast not in _libs:
        build()
        L = C.CDLL(_lib_path(fast))
        L.orc_rhs_acm_3d.argtypes = [C.POINTER(AcmParams), C.c_int, _ip, _dp, _dp, _dp, _dp]
        L.orc_rhs_acm_2d.argtypes = [C.POINTER(AcmParams), C.c_int, _ip, _dp, _dp, _dp, _dp]
        L.orc_get_dt_block.argtypes = [C.POINTER(AcmParams), C.c_int, _ip, _dp, _dp]
        L.orc_get_dt_block.restype = C.c_double
        L.orc_rk_copy_interior.argtypes = [C.c_int, C.c_int, _ip, C.c_int, _dp, _dp]
        L.orc_rk_axpy_interior.argtypes = [C.c_int, C.c_int, _ip, C.c_int, _dp, C.c_double, C.c_double, _dp]
        L.orc_max_abs_interior.argtypes = [C.c_int, C.c_int, _ip, C.c_int, _dp]
        L.orc_max_abs_interior.restype = C.c_double
        L.orc_fd_halfwidth.argtypes = [C.c_int]
        L.orc_sync_same_level.argtypes = [C.c_int, _ip, C.c_int, C.c_int, _ip, C.c_int, _dp, C.c_int, C.c_int]
        L.orc_dt_tree.argtypes = [C.POINTER(AcmParams), C.c_int, _dp, C.c_int, C.c_int, _ip, C.c_int, _dp]
        L.orc_dt_tree.restype = C.c_double
        L.orc_rhs_tree.argtypes = [C.POINTER(AcmParams), C.c_int, _dp, C.c_int, C.c_int, _ip, C.c_int, _dp, _dp, _dp, C.c_int]
        L.orc_rk_step_same_level.argtypes = [C.POINTER(AcmParams), C.c_int, _ip, _dp, C.c_int, C.c_int, C.c_int, _ip, C.c_int,
                                             _dp, _dp, _dp, C.c_int, C.c_double, _dp, C.c_int]
        _libs[fast] = L
    return _libs[fast]


def _p(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_dpp)


def _pb(a: np.ndarray):
    """addresses of a[0], a[1], ... (the per-block loops call C once per block: one .ctypes object per array instead of one per block).
    Plain integers: the caller keeps `a` alive."""
    assert a.dtype == np.float64
    if a.flags.c_contiguous:
        base, st = a.ctypes.data, a.strides[0]
        return [base + b * st for b in range(a.shape[0])]
    out = []
    for b in range(a.shape[0]):                    # e.g. a strided selection of blocks: every block itself must be contiguous
        assert a[b].flags.c_contiguous
        out.append(a[b].ctypes.data)
    return out


@functools.lru_cache(maxsize=None)
def _bs_cached(Bs):
    return (C.c_int32 * 3)(*Bs)


def _bs(Bs: Sequence[int]):
    return _bs_cached(tuple(int(b) for b in Bs))          # read-only in every C routine


def _d3(v: Sequence[float]):
    return (C.c_double * 3)(*[float(x) for x in v])


# ----------------------------------------------------------------------------- configuration
BUTCHER_RK4 = np.array([[0.0, 0.0, 0.0, 0.0, 0.0],
                        [0.5, 0.5, 0.0, 0.0, 0.0],
                        [0.5, 0.0, 0.5, 0.0, 0.0],
                        [1.0, 0.0, 0.0, 1.0, 0.0],
                        [0.0, 1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0]])  # ini_file_to_params.f90:640-646


@dataclass
class Params:
    """Subset of type_params / type_params_acm the hot path reads (module_params.f90:18-233,
    module_ACM.f90:56-154).  Defaults are the reference's read_param defaults."""
    dim: int = 3
    Bs: Tuple[int, int, int] = (16, 16, 16)
    g: int = 3
    g_rhs: int = 2
    n_eqn: int = 4
    domain: Tuple[float, float, float] = (1.0, 1.0, 1.0)
    Jmax: int = 1
    discretization: str = "FD_4th_central"
    skew: bool = False
    penalization: bool = False
    use_sponge: bool = False
    c0: float = 10.0
    nu: float = 1e-1
    gamma_p: float = 1.0
    C_eta: float = 1.0
    C_sponge: float = 1.0e-2
    u_mean_set: Tuple[float, float, float] = (1.0, 0.0, 0.0)
    CFL: float = 1.0
    CFL_eta: float = 0.99
    CFL_nu: Optional[float] = None
    dt_fixed: float = 0.0
    dt_max: float = 0.0
    time_max: float = 1.0
    write_method: str = "fixed_freq"
    write_time: float = 1.0
    write_time_first: float = 0.0
    tsave_stats: float = 9999999.9
    butcher: np.ndarray = field(default_factory=lambda: BUTCHER_RK4.copy())

    def acm(self) -> AcmParams:
        a = AcmParams()
        a.dim, a.fd = self.dim, FD_IDS[self.discretization]
        a.skew, a.penalization, a.use_sponge = int(self.skew), int(self.penalization), int(self.use_sponge)
        a.c0, a.nu, a.gamma_p, a.C_eta, a.C_sponge = self.c0, self.nu, self.gamma_p, self.C_eta, self.C_sponge
        a.u_mean_set = (C.c_double * 3)(*self.u_mean_set)
        a.CFL, a.CFL_eta = self.CFL, self.CFL_eta
        a.CFL_nu = self.cfl_nu()
        return a

    def cfl_nu(self) -> float:
        """module_ACM.f90:367-381 -- default depends on the order digit of the discretization."""
        if self.CFL_nu is not None:
            return self.CFL_nu
        digit = self.discretization[3]
        den = {"2": 4.000, "4": 5.333, "6": 6.0444}[digit]
        return 0.95 * 2.79 / (den * float(self.dim))


# ----------------------------------------------------------------------------- grid
@dataclass
class Grid:
    """Light data of a single tree: integer block coordinates (ix,iy,iz) at `level` per block.

    Block spacing / origin follow module_treelib.f90:93-98:
        dx = 2^-J * L / Bs ;  x0 = (ixyz) * Bs * dx          (ixyz zero-based here)
    """
    level: np.ndarray   # (Nb,) int
    ixyz: np.ndarray    # (Nb,3) int, zero-based block coordinates on their own level
    dim: int = 3

    @property
    def n(self) -> int:
        return len(self.level)

    def spacing_origin(self, p: Params, b: int):
        J = int(self.level[b])
        dx = np.zeros(3)
        x0 = np.zeros(3)
        for d in range(self.dim):
            dx[d] = 2.0 ** (-J) * p.domain[d] / float(p.Bs[d])
            x0[d] = float(int(self.ixyz[b, d]) * p.Bs[d]) * dx[d]
        return x0, dx

    def same_level_neighbors(self) -> Dict[Tuple[int, int, int], np.ndarray]:
        """Periodic same-level neighbour block index per direction (every block must have one)."""
        if getattr(self, "_nbr", None) is None:
            look = self.lookup()
            out = {}
            for dz_ in ((-1, 0, 1) if self.dim == 3 else (0,)):
                for dy_ in (-1, 0, 1):
                    for dx_ in (-1, 0, 1):
                        d = (dx_, dy_, dz_)
                        idx = np.zeros(self.n, dtype=np.int64)
                        for b in range(self.n):
                            J = int(self.level[b])
                            nblk = 2 ** J
                            q = [(int(self.ixyz[b, a]) + d[a]) % nblk if a < self.dim else 0 for a in range(3)]
                            idx[b] = look[(J, q[0], q[1], q[2])]
                        out[d] = idx
            self._nbr = out
        return self._nbr

    def lookup(self) -> Dict[Tuple[int, int, int, int], int]:
        return {(int(l), int(i[0]), int(i[1]), int(i[2])): k for k, (l, i) in enumerate(zip(self.level, self.ixyz))}


def uniform_grid(J: int, dim: int = 3) -> Grid:
    n = 2 ** J
    if dim == 3:
        iz, iy, ix = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    else:
        iy, ix = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
        iz = np.zeros_like(ix)
    ixyz = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1)
    return Grid(level=np.full(len(ixyz), J, dtype=np.int64), ixyz=ixyz.astype(np.int64), dim=dim)


def alloc(grid: Grid, p: Params, nc: Optional[int] = None) -> np.ndarray:
    nc = p.n_eqn if nc is None else nc
    nz = p.Bs[2] + 2 * p.g if p.dim == 3 else 1
    return np.zeros((grid.n, nc, nz, p.Bs[1] + 2 * p.g, p.Bs[0] + 2 * p.g))


def interior(p: Params):
    g = p.g
    zs = slice(g, p.Bs[2] + g) if p.dim == 3 else slice(0, 1)
    return (zs, slice(g, p.Bs[1] + g), slice(g, p.Bs[0] + g))


# ----------------------------------------------------------------------------- ghost sync (same level)
def sync_ghosts_same_level(grid: Grid, p: Params, hvy: np.ndarray, g_minus: int, g_plus: int,
                           ncomp: Optional[int] = None) -> None:
    """Same-level, periodic ghost synchronisation (lvl_diff = 0 relations only).

    Receiver boxes: get_indices_of_ghost_patch (neighborhood.f90:158-285): fixed direction '-':
    g-gminus+1:g, '+': Bs+g+1:Bs+g+gplus, free directions g+1:Bs+g.  Sender boxes: interior
    strip of matching depth (set_send_bounds, calc_data_bounds.f90:99-146).  A copy, so exact.
    All 26 (8 in 2-D) relations are exchanged as in sync_ghosts_generic stage 1
    (synchronize_ghosts_generic.f90:266-339).
    """
    g, Bs, dim = p.g, p.Bs, grid.dim
    nc = hvy.shape[1] if ncomp is None else ncomp
    dirs = [(dx_, dy_, dz_) for dz_ in ((-1, 0, 1) if dim == 3 else (0,)) for dy_ in (-1, 0, 1) for dx_ in (-1, 0, 1)
            if (dx_, dy_, dz_) != (0, 0, 0)]
    src = hvy  # senders are interior strips, receivers ghost strips: disjoint, so in-place is safe

    def boxes(d, n):
        if d == 0:
            return slice(g, n + g), slice(g, n + g)
        if d < 0:   # my lower ghost  <- neighbour's upper interior strip
            return slice(g - g_minus, g), slice(n + g - g_minus, n + g)
        return slice(n + g, n + g + g_plus), slice(g, g + g_plus)

    nbr = grid.same_level_neighbors()
    for d in dirs:
        nb = nbr[d]
        rx, sx = boxes(d[0], Bs[0])
        ry, sy = boxes(d[1], Bs[1])
        if dim == 3:
            rz, sz = boxes(d[2], Bs[2])
        else:
            rz = sz = slice(0, 1)
        hvy[:, :nc, rz, ry, rx] = src[:, :nc, sz, sy, sx][nb]


# ----------------------------------------------------------------------------- RHS / dt / RK
def rhs_tree(grid: Grid, p: Params, hvy: np.ndarray, rhs: np.ndarray, mask: Optional[np.ndarray] = None,
             fast: bool = False) -> None:
    """RHS_wrapper local_stage loop (RHS_wrapper.f90:124-142) -> RHS_ACM -> RHS_{2,3}D_acm."""
    L = lib(fast)
    a = C.byref(p.acm())
    f = L.orc_rhs_acm_3d if p.dim == 3 else L.orc_rhs_acm_2d
    Bs, ph, pr = _bs(p.Bs), _pb(hvy), _pb(rhs)
    pm = None if mask is None else _pb(mask)
    for b in range(grid.n):
        _, dx = grid.spacing_origin(p, b)
        m = None if pm is None else pm[b if len(pm) > 1 else 0]
        f(a, p.g, Bs, _d3(dx), ph[b], pr[b], m)


LIM_DIVERGED = 1.0e12


def divergence_guard(grid: Grid, p: Params, hvy: np.ndarray) -> bool:
    """integral_stage guard (rhs_ACM.f90:133-146): True if any |u| > 1e12 in a block interior."""
    L = lib()
    for b in range(grid.n):
        if L.orc_max_abs_interior(p.dim, p.g, _bs(p.Bs), hvy.shape[1], _p(hvy[b])) > LIM_DIVERGED:
            return True
    return False


def calculate_time_step(grid: Grid, p: Params, hvy: np.ndarray, time: float) -> float:
    """calculate_time_step.f90:19-118 (statistics / probes / backup clauses that are off by default omitted)."""
    dt = 9.0e9
    if p.dt_fixed > 0.0:
        dt = p.dt_fixed
    else:
        L = lib()
        a, Bs, ph = C.byref(p.acm()), _bs(p.Bs), _pb(hvy)
        for b in range(grid.n):
            _, dx = grid.spacing_origin(p, b)
            dt = min(dt, L.orc_get_dt_block(a, p.g, Bs, _d3(dx), ph[b]))
        if p.dt_max > 0.0:
            dt = min(p.dt_max, dt)
    return _clip_dt(p, dt, time, apply_dt_max=False)


def _clip_dt(p: Params, dt: float, time: float, apply_dt_max: bool = True) -> float:
    """calculate_time_step.f90:53-118 -- everything after the global MIN."""
    if apply_dt_max and p.dt_max > 0.0 and not p.dt_fixed > 0.0:
        dt = min(p.dt_max, dt)
    if p.write_method == "fixed_time":
        if (math.fmod(time + dt, p.write_time) < math.fmod(time + 1e-12, p.write_time)
                and not abs(math.fmod(time, p.write_time)) < 1e-12 and time + 1e-12 > p.write_time_first):
            dt = p.write_time - math.fmod(time, p.write_time)
    if abs(p.tsave_stats - 9999999.9) > 1e-3:
        if (math.fmod(time + dt, p.tsave_stats) < math.fmod(time + 1e-12, p.tsave_stats)
                and not abs(math.fmod(time, p.tsave_stats)) < 1e-12):
            dt = p.tsave_stats - math.fmod(time, p.tsave_stats)
    if time + dt > p.time_max and time <= p.time_max:
        dt = p.time_max - time
    return dt


def rk_generic(grid: Grid, p: Params, hvy: np.ndarray, work: np.ndarray, time: float,
               mask: Optional[np.ndarray] = None, sync=None, fast: bool = False, mask_at=None) -> float:
    """RungeKuttaGeneric (runge_kutta_generic.f90:50-154).  `work[slot]` are ghosted arrays like hvy.

    `sync(hvy)` performs sync_ghosts_RHS_tree with g_minus=g_plus=g_rhs; defaults to the same-level sync.
    Returns dt.  hvy is advanced in place.
    """
    L = lib(fast)
    rk = p.butcher
    n = rk.shape[0]
    nc = p.n_eqn
    Bs = _bs(p.Bs)
    if sync is None:
        sync = lambda h: sync_ghosts_same_level(grid, p, h, p.g_rhs, p.g_rhs)
    sync(hvy)
    dt = calculate_time_step(grid, p, hvy, time)
    ph, pw = _pb(hvy), [_pb(w_) for w_ in work]
    for b in range(grid.n):
        L.orc_rk_copy_interior(p.dim, p.g, Bs, nc, pw[0][b], ph[b])
    if mask_at is not None:          # time-dependent mask: RHS_wrapper calls createMask_tree at the stage time (RHS_wrapper.f90:51)
        mask = mask_at(time)
    rhs_tree(grid, p, hvy, work[1], mask, fast)
    for j in range(2, n):          # Fortran j = 2 .. size(rk,1)-1
        for b in range(grid.n):
            L.orc_rk_copy_interior(p.dim, p.g, Bs, nc, ph[b], pw[0][b])
        for l in range(2, j + 1):
            coef = rk[j - 1, l - 1]
            if abs(coef) < 1.0e-8:
                continue
            for b in range(grid.n):
                L.orc_rk_axpy_interior(p.dim, p.g, Bs, nc, ph[b], dt, coef, pw[l - 1][b])
        sync(hvy)
        if mask_at is not None:
            mask = mask_at(time + dt * rk[j - 1, 0])                     # t = time + dt*rk_coeffs(j,1), runge_kutta_generic.f90:122
        rhs_tree(grid, p, hvy, work[j], mask, fast)
    for b in range(grid.n):
        L.orc_rk_copy_interior(p.dim, p.g, Bs, nc, ph[b], pw[0][b])
        for j in range(2, rk.shape[1] + 1):
            coef = rk[n - 1, j - 1]
            if abs(coef) < 1.0e-8:
                continue
            L.orc_rk_axpy_interior(p.dim, p.g, Bs, nc, ph[b], dt, coef, pw[j - 1][b])
    return dt


# ----------------------------------------------------------------------------- initial conditions
def inicond_taylor_green(grid: Grid, p: Params, hvy: np.ndarray) -> None:
    """inicond 'taylor-green-vanRees2011' (inicond_ACM.f90:371-389), set on the whole ghosted block."""
    g = p.g
    for b in range(grid.n):
        x0, dx = grid.spacing_origin(p, b)
        x = (np.arange(p.Bs[0] + 2 * g) - g).astype(np.float64) * dx[0] + x0[0]
        y = (np.arange(p.Bs[1] + 2 * g) - g).astype(np.float64) * dx[1] + x0[1]
        z = (np.arange(p.Bs[2] + 2 * g) - g).astype(np.float64) * dx[2] + x0[2]
        Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
        hvy[b, 0] = np.sin(X) * np.cos(Y) * np.cos(Z)
        hvy[b, 1] = -np.cos(X) * np.sin(Y) * np.cos(Z)
        hvy[b, 2] = 0.0
        hvy[b, 3] = (np.cos(2.0 * X) + np.cos(2.0 * Y)) * (np.cos(2.0 * Z) + 2.0) / 16.0


# ----------------------------------------------------------------------------- C tree loops (orc_tree.c)
def nbr_table(grid: Grid) -> np.ndarray:
    """[n,27] same-level neighbour indices, index (dz+1)*9+(dy+1)*3+(dx+1)."""
    nb = grid.same_level_neighbors()
    out = np.full((grid.n, 27), -1, dtype=np.int32)
    for (dx_, dy_, dz_), idx in nb.items():
        out[:, (dz_ + 1) * 9 + (dy_ + 1) * 3 + (dx_ + 1)] = idx
    return np.ascontiguousarray(out)


def dx_table(grid: Grid, p: Params) -> np.ndarray:
    return np.ascontiguousarray(np.stack([grid.spacing_origin(p, b)[1] for b in range(grid.n)]))


def rk_step_c(grid: Grid, p: Params, hvy: np.ndarray, work: np.ndarray, time: float, nbr: np.ndarray, dxb: np.ndarray,
              fast: bool = False, mask: Optional[np.ndarray] = None) -> float:
    """One RungeKuttaGeneric step with every block loop in C/OpenMP (same arithmetic as rk_generic).
    work: array [s+1, n, nc, nz, ny, nx]."""
    L = lib(fast)
    a = p.acm()
    Bs = _bs(p.Bs)
    nc = p.n_eqn
    ip = lambda x: x.ctypes.data_as(_ip)
    L.orc_sync_same_level(grid.n, ip(nbr), p.dim, p.g, Bs, nc, _p(hvy), p.g_rhs, p.g_rhs)
    if p.dt_fixed > 0.0:
        dt = calculate_time_step(grid, p, hvy, time)
    else:
        dt_cfl = L.orc_dt_tree(C.byref(a), grid.n, _p(dxb), p.dim, p.g, Bs, nc, _p(hvy))
        dt = _clip_dt(p, dt_cfl, time)
    s = p.butcher.shape[0] - 1
    L.orc_rk_step_same_level(C.byref(a), grid.n, ip(nbr), _p(dxb), p.dim, p.g, p.g_rhs, Bs, nc, _p(hvy), _p(work),
                             _p(np.ascontiguousarray(p.butcher)), s, dt, _p(mask), 0 if mask is None else mask.shape[1])
    return dt


# ----------------------------------------------------------------------------- wavelets (orc_wavelet.c)
ORC_FMAX = 12


class Wavelet(C.Structure):
    _fields_ = [("X", C.c_int32), ("Y", C.c_int32),
                ("hd_lo", C.c_int32), ("hd_hi", C.c_int32), ("gd_lo", C.c_int32), ("gd_hi", C.c_int32),
                ("hr_lo", C.c_int32), ("hr_hi", C.c_int32), ("gr_lo", C.c_int32), ("gr_hi", C.c_int32),
                ("g_default", C.c_int32), ("lifted", C.c_int32),
                ("Nscl", C.c_int32), ("Nscr", C.c_int32), ("Nwcl", C.c_int32), ("Nwcr", C.c_int32),
                ("Nreconl", C.c_int32), ("Nreconr", C.c_int32),
                ("HD", C.c_double * (2 * ORC_FMAX + 1)), ("GD", C.c_double * (2 * ORC_FMAX + 1)),
                ("HR", C.c_double * (2 * ORC_FMAX + 1)), ("GR", C.c_double * (2 * ORC_FMAX + 1))]

    def taps(self, name):
        lo, hi = getattr(self, name.lower() + "_lo"), getattr(self, name.lower() + "_hi")
        arr = getattr(self, name)
        return {k: arr[k + ORC_FMAX] for k in range(lo, hi + 1)}


def _wl(fast=False):
    L = lib(fast)
    if not hasattr(L, "_wl_ready"):
        W = C.POINTER(Wavelet)
        L.orc_setup_wavelet.argtypes = [C.c_char_p, W]
        L.orc_fwt_block.argtypes = [W, C.c_int, C.c_int, _ip, C.c_int, _dp, _dp]
        L.orc_iwt_block.argtypes = [W, C.c_int, C.c_int, _ip, C.c_int, _dp, _dp]
        L.orc_threshold_block.argtypes = [C.c_int, C.c_int, _ip, C.c_int, _dp, C.c_int, C.c_int, C.c_int, _ip, _dp, _dp, _dp]
        L.orc_prediction.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp]
        L.orc_refine_block.argtypes = [C.c_int, C.c_int, C.c_int, _ip, C.c_int, _dp, _dp]
        L.orc_block_linfty.argtypes = [C.c_int, C.c_int, _ip, C.c_int, _dp, _dp]
        L._wl_ready = True
    return L


def setup_wavelet(name: str) -> Wavelet:
    w = Wavelet()
    rc = _wl().orc_setup_wavelet(name.encode(), C.byref(w))
    if rc:
        raise ValueError(f"unsupported wavelet {name}")
    return w


EPS_NORMS = {"Linfty": 0, "L1": 1, "L2": 2, "H1": 3}


def fwt_tree(w: Wavelet, p: Params, hvy: np.ndarray, out: np.ndarray) -> None:
    """waveletDecomposition_optimized_block on every block (ghosts of hvy must be synchronised to depth g)."""
    L = _wl()
    for b in range(hvy.shape[0]):
        L.orc_fwt_block(C.byref(w), p.dim, p.g, _bs(p.Bs), hvy.shape[1], _p(hvy[b]), _p(out[b]))


def iwt_tree(w: Wavelet, p: Params, hvy_wd: np.ndarray, out: np.ndarray) -> None:
    L = _wl()
    for b in range(hvy_wd.shape[0]):
        L.orc_iwt_block(C.byref(w), p.dim, p.g, _bs(p.Bs), hvy_wd.shape[1], _p(hvy_wd[b]), _p(out[b]))


def threshold_tree(p: Params, hvy_wd: np.ndarray, level, eps: float, norm=None, eps_norm: str = "Linfty", thresh_comp=None,
                   level_ref: int = 0):
    """threshold_block per block: returns (refinement_status[nb], detail[nb, nc])."""
    L = _wl()
    nb, nc = hvy_wd.shape[:2]
    tc = np.ascontiguousarray(np.ones(nc) if thresh_comp is None else thresh_comp, dtype=np.int32)
    e = np.full(nc, eps, dtype=np.float64)
    nrm = None if norm is None else np.ascontiguousarray(norm, dtype=np.float64)
    det = np.zeros((nb, nc))
    st = np.zeros(nb, dtype=np.int32)
    for b in range(nb):
        st[b] = L.orc_threshold_block(p.dim, p.g, _bs(p.Bs), nc, _p(hvy_wd[b]), int(level[b]), level_ref, EPS_NORMS[eps_norm],
                                      tc.ctypes.data_as(_ip), _p(e), _p(nrm), _p(det[b]))
    return st, det


def norm_linfty_tree(p: Params, hvy: np.ndarray) -> np.ndarray:
    """componentWiseNorm_tree, Linfty; the caller applies `norm <= 1e-9 -> 1` (coarseningIndicator_tree.f90:165-167)."""
    L = _wl()
    nrm = np.zeros(hvy.shape[1])
    for b in range(hvy.shape[0]):
        L.orc_block_linfty(p.dim, p.g, _bs(p.Bs), hvy.shape[1], _p(hvy[b]), _p(nrm))
    return nrm


def prediction(order: int, coarse: np.ndarray) -> np.ndarray:
    """coarse[nz,ny,nx] -> fine[2nz-1, 2ny-1, 2nx-1]"""
    c = np.ascontiguousarray(coarse, dtype=np.float64)
    nz, ny, nx = c.shape
    fine = np.zeros((2 * nz - 1, 2 * ny - 1, 2 * nx - 1))
    _wl().orc_prediction(order, nx, ny, nz, _p(c), _p(fine))
    return fine


def refine_block(order: int, p: Params, mother: np.ndarray) -> np.ndarray:
    """mother[nc,nz,ny,nx] ghosted -> daughters[2^dim, nc, nz, ny, nx] ghosted"""
    m = np.ascontiguousarray(mother, dtype=np.float64)
    d = np.zeros((2 ** p.dim,) + m.shape)
    _wl().orc_refine_block(order, p.dim, p.g, _bs(p.Bs), m.shape[0], _p(m), _p(d))
    return d


# ----------------------------------------------------------------------------- level-jump ghost sync (orc_sync.c)
def same_level_code(d) -> int:
    return _same_level_code(tuple(int(v) for v in d))


@functools.lru_cache(maxsize=None)
def _same_level_code(d) -> int:
    """slot of the same-level neighbour in direction d (find_neighbor, LIB/MESH/find_neighbors.f90:60-95)"""
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


@functools.lru_cache(maxsize=None)
def _relations(dim: int):
    """per direction: (d, slot of the same-level relation, number of finer neighbours, their treecode digits) -- find_neighbor's
    direction bookkeeping, which does not depend on the block"""
    vary = (2, 1, 4)
    out = []
    for dz in ((-1, 0, 1) if dim == 3 else (0,)):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                d = (dx, dy, dz)
                if d == (0, 0, 0):
                    continue
                nfree = 2 ** sum(1 for a in range(dim) if d[a] == 0)
                append = [0, 0, 0, 0]
                apply_free = 1
                for a in range(dim):
                    if d[a] == 0:
                        for k in range(4):
                            append[k] += vary[a] * ((k // apply_free) % 2)
                        apply_free += 1
                    elif d[a] == 1:
                        for k in range(4):
                            append[k] += vary[a]
                out.append((d, same_level_code(d), nfree, tuple(append)))
    return tuple(out)


def neighbor_table168(grid: Grid, Jmax: int, periodic=(1, 1, 1)) -> np.ndarray:
    """hvy_neighbor[168, nb] (1-based block ids, -1 none) of a leaf grid on one rank -- find_neighbor (LIB/MESH/find_neighbors.f90:18-180) for
    all blocks at once: same level, else finer (+112), else coarser (+56).  Lookups by binary search in the sorted (level, ix, iy, iz) keys."""
    dim, n = grid.dim, grid.n
    if n < 64:                                     # a handful of blocks: the block loop is quicker than 26 vector passes
        return neighbor_table168_loop(grid, Jmax, periodic)
    out = np.full((168, n), -1, dtype=np.int32)
    vary = (2, 1, 4)
    J = grid.level.astype(np.int64)
    ix = grid.ixyz.astype(np.int64)
    key = lambda lv, x, y, z: (lv << 60) | (x << 40) | (y << 20) | z          # noqa: E731  (levels < 8, coordinates < 2^20)
    keys = key(J, ix[:, 0], ix[:, 1], ix[:, 2])
    order = np.argsort(keys, kind="stable")
    skeys = keys[order]

    def look(lv, q):
        k = key(lv, q[0], q[1], q[2])
        pos = np.minimum(np.searchsorted(skeys, k), n - 1)
        return np.where(skeys[pos] == k, order[pos], -1)
    N = np.int64(1) << J
    tc_last = np.where(J > 0, sum(vary[a] * (ix[:, a] & 1) for a in range(dim)), 0)
    cols = np.arange(n)
    zero = np.zeros(n, dtype=np.int64)
    for d, code, nfree, append in _relations(dim):
        p = [ix[:, a] + d[a] if a < dim else zero for a in range(3)]
        ok = np.ones(n, dtype=bool)
        for a in range(dim):
            if not periodic[a]:
                ok &= (p[a] >= 0) & (p[a] < N)
        p = [p[a] % N if a < dim else zero for a in range(3)]
        same = np.where(ok, look(J, p), -1)
        hit = same >= 0
        out[code - 1, cols[hit]] = same[hit] + 1
        todo = ok & ~hit
        found = np.zeros(n, dtype=bool)
        alive = todo & (J < Jmax)
        for k in range(nfree):
            if not alive.any():
                break
            q = [((2 * ix[:, a] + (1 if append[k] & vary[a] else 0) + d[a]) % (2 * N)) if a < dim else zero for a in range(3)]
            fin = np.where(alive, look(J + 1, q), -1)
            alive = alive & (fin >= 0)                      # the reference stops at the first missing finer neighbour
            out[code - 1 + k + 112, cols[alive]] = fin[alive] + 1
            found |= alive
        rest = todo & ~found & (J > 0)
        if rest.any():
            crs = look(J - 1, [p[0] >> 1, p[1] >> 1, p[2] >> 1])
            for k in range(nfree):
                m = rest & (tc_last == append[k]) & (crs >= 0)
                out[code - 1 + k + 56, cols[m]] = crs[m] + 1
    return out


def neighbor_table168_loop(grid: Grid, Jmax: int, periodic=(1, 1, 1)) -> np.ndarray:
    """hvy_neighbor[168, nb] (1-based block ids, -1 none) of a leaf grid on one rank -- an independent restatement of find_neighbor
    (LIB/MESH/find_neighbors.f90:18-180), block by block: same level, else finer (+112), else coarser (+56).  neighbor_table168 is the same
    search for all blocks at once (tests/test_oracle_tree.py compares the two)."""
    dim = grid.dim
    look = grid.lookup()
    out = np.full((168, grid.n), -1, dtype=np.int32)
    vary = (2, 1, 4)
    rel = _relations(dim)
    levels, ixyz = grid.level.tolist(), grid.ixyz.tolist()
    for b in range(grid.n):
        J, ix = levels[b], ixyz[b]
        tc_last = sum(vary[a] * (ix[a] & 1) for a in range(dim)) if J > 0 else 0
        n = 2 ** J
        for d, code, nfree, append in rel:
            p = [ix[a] + d[a] for a in range(3)]
            if any((p[a] < 0 or p[a] >= n) and not periodic[a] for a in range(dim)):
                continue
            p = [p[a] % n if a < dim else 0 for a in range(3)]
            j = look.get((J, p[0], p[1], p[2]))
            if j is not None:
                out[code - 1, b] = j + 1
                continue
            found = False
            if J < Jmax:
                for k in range(nfree):
                    q = [((2 * ix[a] + (1 if append[k] & vary[a] else 0) + d[a]) % (2 * n)) if a < dim else 0 for a in range(3)]
                    j = look.get((J + 1, q[0], q[1], q[2]))
                    if j is None:
                        break
                    out[code - 1 + k + 112, b] = j + 1
                    found = True
            if found:
                continue
            for k in range(nfree):
                if tc_last == append[k] and J > 0:
                    j = look.get((J - 1, p[0] >> 1, p[1] >> 1, p[2] >> 1))
                    if j is not None:
                        out[code - 1 + k + 56, b] = j + 1
    return out


def sync_ghosts_leaf(grid: Grid, p: Params, hvy: np.ndarray, nbr168: np.ndarray, g_minus: int, g_plus: int, order: int,
                     lifted: bool, ignore_filter: bool = True, w: Optional["Wavelet"] = None) -> int:
    """sync_ghosts_generic("full_leaf") on a leaf grid with level jumps (orc_sync.c).  ignore_filter=True is sync_ghosts_RHS_tree
    (plain decimation); ignore_filter=False with a lifted wavelet `w` is sync_ghosts_tree's default: the restriction takes its values
    from the HD-filtered sender (restrict_copy_at_CE, LIB/MPI/restrict_predict_data.f90:121-172)."""
    L = lib()
    if not hasattr(L, "_sync_ready"):
        L.orc_sync_ghosts_leaf_ex.argtypes = [C.c_int, _ip, _ip, C.c_int, C.c_int, _ip, C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_sync_ghosts_leaf_ex.restype = C.c_int
        L.orc_inverse_relation.argtypes = [C.c_int]
        L.orc_inverse_relation.restype = C.c_int
        L._sync_ready = True
    nb = np.ascontiguousarray(nbr168, dtype=np.int32)
    lv = np.ascontiguousarray(grid.level, dtype=np.int32)
    if ignore_filter or not lifted:
        return L.orc_sync_ghosts_leaf_ex(grid.n, nb.ctypes.data_as(_ip), lv.ctypes.data_as(_ip), grid.dim, p.g, _bs(p.Bs), hvy.shape[1],
                                         _p(hvy), g_minus, g_plus, order, int(lifted), 1, None, 0, 0, 0, 0)
    assert w is not None and w.lifted
    hd = np.array(list(w.HD), dtype=np.float64)
    return L.orc_sync_ghosts_leaf_ex(grid.n, nb.ctypes.data_as(_ip), lv.ctypes.data_as(_ip), grid.dim, p.g, _bs(p.Bs), hvy.shape[1], _p(hvy),
                                     g_minus, g_plus, order, 1, 0, _p(hd), w.hd_lo, w.hd_hi, w.Nscl, w.Nscr)


def block_filter(p: Params, u: np.ndarray, coef: dict, do_restriction: bool = False) -> np.ndarray:
    """blockFilterXYZ_vct (module_wavelets.f90:307-401) on one ghosted block u[nc, nz, ny, nx]; coef = {tap: value}."""
    L = lib()
    if not hasattr(L, "_bf_ready"):
        L.orc_block_filter.argtypes = [C.c_int, C.c_int, _ip, C.c_int, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int]
        L.orc_block_filter.restype = None
        L._bf_ready = True
    lo, hi = min(coef), max(coef)
    cf = np.zeros(2 * ORC_FMAX + 1)
    for k, v in coef.items():
        cf[k + ORC_FMAX] = v
    u = np.ascontiguousarray(u, dtype=np.float64)
    out = np.zeros_like(u)
    L.orc_block_filter(p.dim, p.g, _bs(p.Bs), u.shape[0], _p(u), _p(out), _p(cf), lo, hi, int(do_restriction))
    return out


def coarse_extension_modify(grid: Grid, p: Params, w: Wavelet, wd: np.ndarray, orig: np.ndarray, nbr168: np.ndarray, fd_half_width: int = 0,
                            clear_wc: bool = True, copy_sc: bool = True) -> int:
    """coarse_extension_modify("tree") on a leaf grid (LIB/MPI/reconstruction_step.f90:3-100): for every block and every relation
    whose neighbour is coarser, zero the wavelet coefficients / copy the scaling coefficients near the interface.  Nwc includes the
    widening to 2*FD_max_size of setup_wavelet (module_wavelets.f90:1404-1417).  Returns the number of patches treated."""
    L = lib()
    if not hasattr(L, "_ce_ready"):
        L.orc_ce_modify_block.argtypes = [C.c_int, C.c_int, _ip, C.c_int, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L._ce_ready = True
    Nwcl = max(w.Nwcl, 2 * fd_half_width)
    Nwcr = max(w.Nwcr, 2 * fd_half_width)
    n = 0
    for b in range(grid.n):
        for r in range(57, 113):
            if nbr168[r - 1, b] >= 1:
                L.orc_ce_modify_block(p.dim, p.g, _bs(p.Bs), wd.shape[1], _p(wd[b]), _p(orig[b]), r, Nwcl, Nwcr, w.Nscl, w.Nscr,
                                      int(clear_wc), int(copy_sc))
                n += 1
    return n


def rkc_step(grid: Grid, p: Params, hvy: np.ndarray, time: float, mu, mu_tilde, nu, gamma_tilde, c, mask: Optional[np.ndarray] = None, sync=None) -> float:
    """RungeKuttaChebychev (LIB/TIME/runge_kutta_chebychev.f90:6-146), numpy on the whole ghosted arrays (interiors are what counts; ghost nodes
    are re-synchronised before every right-hand side).  mu .. c: rows s of the coefficient tables (setup_RKC_coefficients, :180 ff, or the
    RKC_custom_scheme of the parameter file), 0-based here.  Evaluation order of the main formula as the Fortran expression: left to right.
    hvy is advanced in place; returns dt."""
    s = len(mu)
    if s < 4:
        raise ValueError("runge-kutta-chebychev: s cannot be less than 4")          # abort(1715929)
    if sync is None:
        sync = lambda h: sync_ghosts_same_level(grid, p, h, p.g_rhs, p.g_rhs)
    sync(hvy)
    dt = calculate_time_step(grid, p, hvy, time)
    y00 = hvy.copy()
    y0 = hvy.copy()
    F0 = np.zeros_like(hvy)
    rhs_tree(grid, p, hvy, F0, mask)
    y1 = y0 + mu_tilde[0] * dt * F0
    y2 = y1
    for i in range(1, s):                       # Fortran i = 2 .. s
        sync(y1)
        F1 = np.zeros_like(hvy)
        rhs_tree(grid, p, y1, F1, mask)
        y2 = (1.0 - mu[i] - nu[i]) * y00 + mu[i] * y1 + nu[i] * y0 + mu_tilde[i] * dt * F1 + gamma_tilde[i] * dt * F0
        if i < s - 1:
            y0, y1 = y1, y2
    hvy[:] = y2
    return dt


def expm_pade(H: np.ndarray, ideg: int = 6) -> np.ndarray:
    """exp(H) as expM_pade computes it (LIB/TIME/krylov.f90:193-226 -> DGPADM of Expokit, R. Sidje, ACM TOMS 24 (1998), restated from the
    published algorithm): scaling by 2^ns with ns = max(0, int(log2 |H|_inf) + 2), the irreducible (ideg, ideg) Pade fraction
    exp(A) ~ I + 2 (E - O)^-1 O with E / O the even / odd part of sum_k c_k A^k, c_k = c_(k-1) (ideg + 1 - k) / (k (2 ideg + 1 - k)),
    ns squarings."""
    H = np.asarray(H, dtype=np.float64)
    m = H.shape[0]
    hnorm = float(np.abs(H).sum(axis=1).max())
    if hnorm == 0.0:
        return np.eye(m)
    ns = max(0, int(math.log(hnorm) / math.log(2.0)) + 2)
    A = H * (1.0 / 2.0 ** ns)
    c = [1.0]
    for k in range(1, ideg + 1):
        c.append(c[-1] * float(ideg + 1 - k) / float(k * (2 * ideg + 1 - k)))
    A2 = A @ A
    I = np.eye(m)
    # Horner in A^2: even part c0 + c2 A^2 + ..., odd part (c1 + c3 A^2 + ...) A
    ev = c[ideg - (ideg % 2)] * I
    for k in range(ideg - (ideg % 2) - 2, -1, -2):
        ev = ev @ A2 + c[k] * I
    od = c[ideg - 1 + (ideg % 2)] * I
    for k in range(ideg - 3 + (ideg % 2), 0, -2):
        od = od @ A2 + c[k] * I
    od = od @ A
    E = I + 2.0 * np.linalg.solve(ev - od, od)
    for _ in range(ns):
        E = E @ E
    return E


def krylov_step(grid: Grid, p: Params, hvy: np.ndarray, time: float, M_max: int = 12, dynamic: bool = False, err_threshold: float = 1.0e-3,
                mask: Optional[np.ndarray] = None, sync=None, dot=None):
    """krylov_time_stepper (LIB/TIME/krylov.f90:1-190): the exponential integrator u(t+dt) = u + dt phi_1(dt J) F(u) on the Krylov space of
    the Jacobian J (finite differences of the right-hand side with eps = |u| sqrt(epsilon)), Arnoldi with modified Gram-Schmidt on the block
    interiors (wabbit_norm, scalarproduct :500-595), phi_1 from the matrix exponential of the augmented Hessenberg matrix, error estimate
    |beta h_(M+1,M) phi(M, M+2)|; "dynamic": stop at the first M with err <= threshold, and at M_max shrink dt by 0.9 until it is.
    hvy is advanced in place; returns (dt, M_iter, err)."""
    if sync is None:
        sync = lambda h: sync_ghosts_same_level(grid, p, h, p.g_rhs, p.g_rhs)
    I = (slice(None), slice(None)) + interior(p)
    if dot is None:                                # scalarproduct; a caller may pass another summation order (see tests/test_oracle_krylov.py)
        dot = lambda a, b: float((a[I] * b[I]).sum())
    epsm = float(np.finfo(np.float64).eps)
    sync(hvy)
    dt = calculate_time_step(grid, p, hvy, time)
    normv = math.sqrt(dot(hvy, hvy))
    if normv < epsm:
        normv = 1.0
    eps = normv * math.sqrt(epsm)
    R = np.zeros_like(hvy)
    rhs_tree(grid, p, hvy, R, mask)
    beta = math.sqrt(dot(R, R))
    if beta < epsm:
        beta = 1.0
    V = [R / beta]
    H = np.zeros((M_max + 2, M_max + 2))
    phi = np.zeros((M_max + 2, M_max + 2))
    err, M_iter = 0.0, 0
    for M_iter in range(1, M_max + 1):
        P = hvy + eps * V[M_iter - 1]
        sync(P)
        W = np.zeros_like(hvy)
        rhs_tree(grid, p, P, W, mask)
        W = (W - R) / eps
        for it in range(1, M_iter + 1):
            H[it - 1, M_iter - 1] = dot(V[it - 1], W)
            W = W - H[it - 1, M_iter - 1] * V[it - 1]
        H[M_iter, M_iter - 1] = math.sqrt(dot(W, W))
        V.append(W / H[M_iter, M_iter - 1])
        if dynamic or M_iter == M_max:
            h_klein = H[M_iter, M_iter - 1]
            Ht = np.zeros((M_iter + 2, M_iter + 2))
            Ht[:M_iter, :M_iter] = H[:M_iter, :M_iter]
            Ht[0, M_iter] = 1.0
            Ht[M_iter, M_iter + 1] = 1.0
            phi[:] = 0.0
            phi[:M_iter + 2, :M_iter + 2] = expm_pade(dt * Ht)
            phi[M_iter, M_iter] = h_klein * phi[M_iter - 1, M_iter + 1]
            err = abs(beta * phi[M_iter, M_iter])
            if dynamic and M_iter == M_max and err > err_threshold:
                while err > err_threshold:
                    dt = 0.90 * dt
                    phi[:] = 0.0
                    phi[:M_iter + 2, :M_iter + 2] = expm_pade(dt * Ht)
                    phi[M_iter, M_iter] = h_klein * phi[M_iter - 1, M_iter + 1]
                    err = abs(beta * phi[M_iter, M_iter])
            if err <= err_threshold or M_iter == M_max:
                break
    for it in range(1, M_iter + 2):
        hvy[:] = hvy + beta * V[it - 1] * phi[it - 1, M_iter]
    return dt, M_iter, err


def superviscosity_stencil(filter_type: str) -> dict:
    """The stencil filter_wrapper applies (LIB/TIME/filter_wrapper.f90:28-62, generate_superviscosity_stencil :82-104): binomial coefficients with
    alternating sign, normalised by the sum of their absolute values, negated for explicit_5pt / 9pt / 13pt / 17pt / 21pt, plus the identity.
    {shift: coefficient}."""
    import re
    m = re.match(r"explicit_(\d+)pt$", filter_type)
    if m:
        order = int(m.group(1)) - 1
    else:
        m = re.match(r"superviscosity_(\d+)(?:nd|th)$", filter_type)
        if not m:
            raise ValueError("ERROR: Filter not known: " + filter_type)          # abort(251107)
        order = int(m.group(1))
    if order < 2 or order > 20 or order % 2:
        raise ValueError("ERROR: Filter not known: " + filter_type)
    a = order // 2
    st = np.array([(-1.0) ** (k + a) * float(math.comb(2 * a, a + k)) for k in range(-a, a + 1)])
    st = st / np.abs(st).sum()
    if a % 2 == 0:
        st = -st
    st[a] = st[a] + 1.0
    return {k: float(st[k + a]) for k in range(-a, a + 1)}


def filter_wrapper(grid: Grid, p: Params, hvy: np.ndarray, filter_type: str, filter_component=None, only_maxlevel: bool = False,
                   all_except_maxlevel: bool = False) -> None:
    """filter_wrapper (LIB/TIME/filter_wrapper.f90:1-78) on ghost-synchronised data, in place: blockFilterXYZ_vct per selected block and component"""
    if only_maxlevel and all_except_maxlevel:
        raise ValueError("251106")
    coef = superviscosity_stencil(filter_type)
    if max(coef) > p.g:
        raise ValueError("251108")
    I = interior(p)
    for b in range(grid.n):
        lvl = int(grid.level[b])
        if (only_maxlevel and lvl < p.Jmax) or (all_except_maxlevel and lvl == p.Jmax):
            continue
        for c in range(hvy.shape[1]):
            if filter_component is not None and not filter_component[c]:
                continue
            out = block_filter(p, hvy[b, c:c + 1], coef)
            hvy[b, c][I] = out[0][I]


FD1 = {"FD_2nd_central": (1, [-0.5, 0.0, 0.5]), "FD_4th_central": (2, [1.0 / 12.0, -2.0 / 3.0, 0.0, 2.0 / 3.0, -1.0 / 12.0]),
       "FD_6th_central": (3, [-1.0 / 60.0, 3.0 / 20.0, -3.0 / 4.0, 0.0, 3.0 / 4.0, -3.0 / 20.0, 1.0 / 60.0])}


FD2 = {"FD_2nd_central": (1, [1.0, -2.0, 1.0]), "FD_4th_central": (2, [-1.0 / 12.0, 16.0 / 12.0, -30.0 / 12.0, 16.0 / 12.0, -1.0 / 12.0]),
       "FD_6th_central": (3, [2.0 / 180.0, -27.0 / 180.0, 270.0 / 180.0, -490.0 / 180.0, 270.0 / 180.0, -27.0 / 180.0, 2.0 / 180.0])}
# FD1_C2/C4/C6, FD2_C2/C4/C6 of LIB/OPERATORS/module_operators.f90:25-33


def _fd_sum(comp: np.ndarray, I, ax: int, H: int, coef) -> np.ndarray:
    """sum(FD(s:e) * u(i+s:i+e)) on the interior along array axis ax: every tap (the zero centre included), in increasing tap order from 0 --
    the Fortran SUM of the array product (compute_vorticity.f90:40-45, compute_dissipation.f90:41-69)"""
    acc = np.zeros_like(comp[I])
    for k, c in enumerate(coef):
        sl = list(I)
        s0 = sl[ax]
        sl[ax] = slice(s0.start + k - H, s0.stop + k - H)
        acc = acc + c * comp[tuple(sl)]
    return acc


def vorticity_block(p: Params, u: np.ndarray, dx) -> list:
    """compute_vorticity (LIB/OPERATORS/compute_vorticity.f90:3-67) on the interior of one ghost-synchronised block u[c, z, y, x]:
    [v_dx - u_dy] in 2-D, [w_dy - v_dz, u_dz - w_dx, v_dx - u_dy] in 3-D.  Pinned bit for bit by the vor / vorabs files the reference saved
    next to its regression fields (tests/test_oracle_derived_fields.py)."""
    I = interior(p)
    H1, a1 = FD1[p.discretization]
    AX = {0: 2, 1: 1, 2: 0}                        # u[c, z, y, x]: x is the last axis
    d1 = lambda c, d: _fd_sum(u[c], I, AX[d], H1, a1) * (1.0 / dx[d])          # noqa: E731
    if p.dim == 2:
        return [d1(1, 0) - d1(0, 1)]
    return [d1(2, 1) - d1(1, 2), d1(0, 2) - d1(2, 0), d1(1, 0) - d1(0, 1)]


def divergence_block(p: Params, u: np.ndarray, dx) -> np.ndarray:
    """divergence (LIB/OPERATORS/divergence.f90) on the interior of one ghost-synchronised block: u_dx + v_dy (+ w_dz), the first-derivative
    stencil of the discretization; pinned bit for bit by the div files of the reference's 3vortices cases"""
    I = interior(p)
    H1, a1 = FD1[p.discretization]
    AX = {0: 2, 1: 1, 2: 0}
    out = None
    for d in range(p.dim):
        t = _fd_sum(u[d], I, AX[d], H1, a1) * (1.0 / dx[d])
        out = t if out is None else out + t
    return out


def vorticity_statistics_acm(grid: Grid, p: Params, hvy: np.ndarray) -> dict:
    """enstrophy, max_vort, helicity and dissipation of STATISTICS_ACM (statistics_ACM.f90:371-387): compute_vorticity (LIB/OPERATORS/
    compute_vorticity.f90:3-67) and compute_dissipation (compute_dissipation.f90:5-78) on the ghost-synchronised state with the module's
    first- and second-derivative stencils, block sums times dV, over the whole domain (penalized regions included).
    hvy: [nb, nc, nz, ny, nx]."""
    dim = p.dim
    I = interior(p)
    H1, a1 = FD1[p.discretization]
    H2, a2 = FD2[p.discretization]
    AX = {0: 2, 1: 1, 2: 0}                        # u[c, z, y, x]: x is the last axis
    out = {"enstrophy": 0.0, "max_vort": 0.0, "helicity": 0.0, "dissipation": 0.0}
    for b in range(grid.n):
        dx = [2.0 ** (-float(grid.level[b])) * p.domain[d] / float(p.Bs[d]) for d in range(dim)]
        dV = float(np.prod(dx))
        u = hvy[b]
        vor = vorticity_block(p, u, dx)
        if dim == 2:
            out["max_vort"] = max(out["max_vort"], float(np.abs(vor[0]).max()))
        else:
            out["max_vort"] = max(out["max_vort"], float(np.sqrt(vor[0] ** 2 + vor[1] ** 2 + vor[2] ** 2).max()))
            out["helicity"] += 0.5 * float(sum((vor[c] * u[c][I]).sum() for c in range(3))) * dV
        out["enstrophy"] += 0.5 * float(sum((v * v).sum() for v in vor)) * dV
        if p.nu > 0.0:
            eps = np.zeros_like(u[0][I])
            for c in range(dim):
                lap = None
                for d in range(dim):
                    t = _fd_sum(u[c], I, AX[d], H2, a2) * (1.0 / dx[d] ** 2)
                    lap = t if lap is None else lap + t
                eps = eps + u[c][I] * lap
            out["dissipation"] -= p.nu * float(eps.sum()) * dV
    return out


def statistics_acm(grid: Grid, p: Params, hvy: np.ndarray, mask: Optional[np.ndarray] = None) -> dict:
    """The integral_stage of STATISTICS_ACM (LIB/EQUATION/ACMnew/statistics_ACM.f90:138-368) summed over the blocks (post_stage, :396-430), numpy.
    hvy: ghost-synchronised state [nb, nc, nz, ny, nx]; mask: hvy_mask [nb, 6, nz, ny, nx] or None.  compute_divergence: central differences of
    the module's discretization (LIB/OPERATORS/divergence.f90), set to zero where mask(1) > 0."""
    dim, g = p.dim, p.g
    I = interior(p)
    H, a = FD1[p.discretization]
    out = {k: 0.0 for k in ("meanflow_x", "meanflow_y", "meanflow_z", "e_kin", "ACM_energy", "mask_volume", "sponge_volume", "penal_power_solid_input",
                            "penal_power_solid_dissipation", "penal_power_sponge", "force_x", "force_y", "force_z", "umag", "div_max", "div_min",
                            "u_residual_x", "u_residual_y", "u_residual_z")}
    C_eta_inv, C_sp_inv = 1.0 / p.C_eta, 1.0 / p.C_sponge
    names = "xyz"
    for b in range(grid.n):
        dx = [2.0 ** (-float(grid.level[b])) * p.domain[d] / float(p.Bs[d]) for d in range(dim)]
        dV = float(np.prod(dx))
        u = hvy[b]
        vel = [u[d][I] for d in range(dim)]
        pr = u[dim][I]
        for d in range(dim):
            out["meanflow_" + names[d]] += float(vel[d].sum()) * dV
        ek = 0.5 * sum(float((v * v).sum()) for v in vel)
        out["e_kin"] += ek * dV
        out["ACM_energy"] += (0.5 * float((pr * pr).sum()) / p.c0 ** 2 + ek) * dV
        out["umag"] = max(out["umag"], float(sum(v * v for v in vel).max()))
        div = divergence_block(p, u, dx)
        if mask is not None:
            chi = mask[b][0][I]
            div = np.where(chi > 0.0, 0.0, div)
        out["div_max"] = max(out["div_max"], float(div.max()))
        out["div_min"] = min(out["div_min"], float(div.min()))
        if mask is not None and (p.penalization or p.use_sponge):
            us = [mask[b][1 + d][I] for d in range(dim)]
            sp = mask[b][5][I] if (p.use_sponge and mask.shape[1] > 5) else np.zeros_like(chi)
            out["mask_volume"] += float(chi.sum()) * dV
            out["sponge_volume"] += float(sp.sum()) * dV
            out["penal_power_solid_input"] += float((sum(us[d] * (vel[d] - us[d]) for d in range(dim)) * chi * C_eta_inv).sum()) * dV
            out["penal_power_solid_dissipation"] += float((sum((vel[d] - us[d]) ** 2 for d in range(dim)) * chi * C_eta_inv).sum()) * dV
            if p.use_sponge:
                out["penal_power_sponge"] += float(((sum(vel[d] * (vel[d] - p.u_mean_set[d]) for d in range(dim)) + pr * pr / p.c0 ** 2) * sp * C_sp_inv).sum()) * dV
            for d in range(dim):
                out["force_" + names[d]] += float((chi * (vel[d] - us[d]) * C_eta_inv).sum()) * dV
                out["u_residual_" + names[d]] += float((np.abs(vel[d] - us[d]) * chi).max()) * dV
    return out
