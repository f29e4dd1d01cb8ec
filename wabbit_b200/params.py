"""WABBIT `.ini` reader and the parameter set of the block hot path.

Keeps the reference's file format and key names (LIB/PARAMS/module_ini_files_parser.f90; keys read in
LIB/MESH/ini_file_to_params.f90:2-647 and LIB/EQUATION/ACMnew/module_ACM.f90:172-609) and the reference's
defaults, so the same PARAMS.ini drives both codes.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from ._native import WGPU_MAX_STAGES, WgpuConfig

FD_IDS = {"FD_2nd_central": 2, "FD_4th_central": 4, "FD_6th_central": 6, "FD_4th_central_optimized": 40}
FD_HALFWIDTH = {2: 1, 4: 2, 6: 3, 40: 3}

BUTCHER_RK4 = [[0.0, 0.0, 0.0, 0.0, 0.0],
               [0.5, 0.5, 0.0, 0.0, 0.0],
               [0.5, 0.0, 0.5, 0.0, 0.0],
               [1.0, 0.0, 0.0, 1.0, 0.0],
               [0.0, 1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0]]   # ini_file_to_params.f90:640-646


class IniFile:
    """`[Section]` / `key=value;` files; everything after the first `;` of a value is a comment; lines starting
    with `;`, `!`, `#` are comments.  An empty value (`key=;`) means "use the default", as in read_param_mpi."""

    def __init__(self, path: str):
        self.sections: Dict[str, Dict[str, str]] = {}
        sec = None
        matrix_key = None
        with open(path, "r") as f:
            for raw in f:
                line = raw.strip()
                if not line or line[0] in ";!#%":
                    continue
                m = re.match(r"^\[(.+?)\]", line)
                if m:
                    sec = m.group(1)
                    self.sections.setdefault(sec, {})
                    matrix_key = None
                    continue
                if sec is None:
                    continue
                if matrix_key is not None:   # continuation rows of a (/ ... /) matrix
                    self.sections[sec][matrix_key] += " " + line.split(";")[0]
                    if "/)" in line:
                        matrix_key = None
                    continue
                if "=" not in line:
                    continue
                key, val = line.split("=", 1)
                key = key.strip()
                if val.strip().startswith("(/"):
                    self.sections[sec][key] = val.split(";")[0]
                    if "/)" not in val:
                        matrix_key = key
                    continue
                self.sections[sec][key] = val.split(";")[0].strip()

    def get(self, sec: str, key: str) -> Optional[str]:
        v = self.sections.get(sec, {}).get(key)
        return v if v not in (None, "") else None

    def real(self, sec, key, default):
        v = self.get(sec, key)
        return float(v.replace("d", "e").replace("D", "e")) if v is not None else default

    def integer(self, sec, key, default):
        v = self.get(sec, key)
        return int(v) if v is not None else default

    def boolean(self, sec, key, default):
        v = self.get(sec, key)
        if v is None:
            return default
        return v.strip().lower() in ("1", "yes", "true", ".true.", "t", "y")

    def string(self, sec, key, default):
        v = self.get(sec, key)
        return v if v is not None else default

    def vector(self, sec, key, default, typ=float):
        v = self.get(sec, key)
        if v is None:
            return default
        return [typ(x) for x in v.replace(",", " ").split()]

    def matrix(self, sec, key, default):
        v = self.get(sec, key)
        if v is None:
            return default
        body = v.replace("(/", "").replace("/)", "")
        rows = [[float(x) for x in r.replace(",", " ").split()] for r in re.split(r"[\n]|  +", body) if r.strip()]
        flat = [x for r in rows for x in r]
        n = int(round(len(flat) ** 0.5))
        return [flat[i * n:(i + 1) * n] for i in range(n)]


def wavelet_ghosts(wavelet: str) -> Tuple[int, int, int, int]:
    """CDFXY -> (X, Y, g_default, g_rhs_default)   (ini_file_to_params.f90:440-475)."""
    m = re.match(r"^CDF(\d)(\d)$", wavelet)
    if not m:
        raise ValueError(f"unsupported wavelet {wavelet!r}")
    X, Y = int(m.group(1)), int(m.group(2))
    return X, Y, X - 1 + max(Y - 1, 0), X // 2


def rkc2_coefficients(s: int, eps: float = 10.0):
    """mu, mu_tilde, nu, gamma_tilde, c (each of length s, entry j-1 for stage j) of the damped second-order Runge-Kutta-Chebychev scheme:
    w0 = 1 + eps / s^2, w1 = T_s'(w0) / T_s''(w0), b_j = T_j''(w0) / T_j'(w0)^2 (b_0 = b_2, b_1 = 1 / w0), a_j = 1 - b_j T_j(w0),
    mu_j = 2 b_j w0 / b_(j-1), nu_j = -b_j / b_(j-2), mu~_1 = b_1 w1, mu~_j = 2 b_j w1 / b_(j-1), gamma~_j = -a_(j-1) mu~_j,
    c_1 = mu~_1, c_j = w1 T_j''(w0) / T_j'(w0)."""
    if s < 2:
        raise ValueError("runge-kutta-chebychev: s must be at least 2")
    w0 = 1.0 + eps / float(s) ** 2
    T, dT, ddT = [1.0, w0], [0.0, 1.0], [0.0, 0.0]
    for k in range(2, s + 1):
        T.append(2.0 * w0 * T[k - 1] - T[k - 2])
        dT.append(2.0 * T[k - 1] + 2.0 * w0 * dT[k - 1] - dT[k - 2])
        ddT.append(4.0 * dT[k - 1] + 2.0 * w0 * ddT[k - 1] - ddT[k - 2])
    w1 = dT[s] / ddT[s]
    b = [0.0] * (s + 1)
    for j in range(2, s + 1):
        b[j] = ddT[j] / dT[j] ** 2
    b[0], b[1] = b[2], 1.0 / w0
    a = [1.0 - b[j] * T[j] for j in range(s + 1)]
    mu, mut, nu, gt, c = (np.zeros(s) for _ in range(5))
    mut[0] = b[1] * w1
    c[0] = mut[0]
    for j in range(2, s + 1):
        mu[j - 1] = 2.0 * b[j] * w0 / b[j - 1]
        nu[j - 1] = -b[j] / b[j - 2]
        mut[j - 1] = 2.0 * b[j] * w1 / b[j - 1]
        gt[j - 1] = -a[j - 1] * mut[j - 1]
        c[j - 1] = w1 * ddT[j] / dT[j]
    return mu, mut, nu, gt, c


@dataclass
class Params:
    dim: int = 3
    domain: Tuple[float, float, float] = (1.0, 1.0, 1.0)
    periodic: Tuple[int, int, int] = (1, 1, 1)
    wavelet: str = "CDF40"
    Bs: Tuple[int, int, int] = (16, 16, 16)
    g: int = 3
    g_rhs: int = 2
    n_eqn: int = 4
    Jmax: int = 5
    Jmin: int = 1
    max_blocks: int = 0
    eps: float = 1e-3
    eps_normalized: bool = False
    eps_norm: str = "Linfty"
    force_maxlevel_dealiasing: bool = False
    useCoarseExtension: int = -1          # -1: the reference's default = isLiftedWavelet (ini_file_to_params.f90:543-546)
    useSecurityZone: int = -1
    adapt_tree: bool = False
    refinement_indicator: str = "everywhere"
    threshold_state_vector_component: Tuple[int, ...] = ()      # () = all components with their own norm (ini_file_to_params.f90:548-552)
    threshold_mask: bool = False
    adapt_inicond: bool = False
    read_from_files: bool = False
    input_files: Tuple[str, ...] = ()
    block_dist: str = "sfc_hilbert"
    discretization: str = "FD_4th_central"
    time_max: float = 1.0
    nt: int = 99999999
    CFL: float = 1.0
    CFL_eta: float = 0.99
    CFL_nu: float = 0.0
    dt_fixed: float = 0.0
    dt_max: float = 0.0
    write_method: str = "fixed_freq"
    write_time: float = 1.0
    write_time_first: float = 0.0
    write_freq: int = 25                                  # [Time] write_freq (write_method = fixed_freq)
    tsave_stats: float = 9999999.9
    nsave_stats: int = 99999999                           # [Statistics] nsave_stats (ini_file_to_params.f90:188)
    butcher: List[List[float]] = field(default_factory=lambda: [r[:] for r in BUTCHER_RK4])
    time_step_method: str = "RungeKuttaGeneric"          # or "RungeKuttaChebychev" (timeStep_tree.f90:26-58)
    rkc_s: int = 4                                        # [Time] s: stages of the Chebychev scheme
    RKC_custom_scheme: bool = False                       # [Time] RKC_custom_scheme: the coefficient rows come from the parameter file
    RKC_mu: Tuple[float, ...] = ()
    RKC_mu_tilde: Tuple[float, ...] = ()
    RKC_nu: Tuple[float, ...] = ()
    RKC_gamma_tilde: Tuple[float, ...] = ()
    RKC_c: Tuple[float, ...] = ()
    M_krylov: int = 12                                    # [Time] M_krylov, krylov_err_threshold, krylov_subspace_dimension (module_params.f90:23-34)
    krylov_err_threshold: float = 1.0e-3
    krylov_subspace_dimension: str = "fixed"
    filter_type: str = "no_filter"                        # [Discretization] filter_type (filter_wrapper.f90)
    filter_freq: int = -1
    filter_only_maxlevel: bool = False
    filter_all_except_maxlevel: bool = False
    filter_component: Tuple[int, ...] = ()                # () = every component
    c0: float = 10.0
    nu: float = 1e-1
    gamma_p: float = 1.0
    u_mean_set: Tuple[float, float, float] = (1.0, 0.0, 0.0)
    skew_symmetry: bool = False
    inicond: str = "meanflow"
    penalization: bool = False
    C_eta: float = 1.0
    use_sponge: bool = False
    C_sponge: float = 1.0e-2
    n_mask: int = 0

    def finalize(self) -> "Params":
        """Derived values and the reference's consistency rules."""
        fd = FD_IDS[self.discretization]
        h = FD_HALFWIDTH[fd]
        for d in range(self.dim):
            if self.Bs[d] % 2:
                raise ValueError("Blocksize Bs must be even")   # module_ini_files_parser_mpi.f90:816-818
        if self.g < h:
            self.g = h                                           # ini_file_to_params.f90:292-298
        if self.g_rhs < h:
            self.g_rhs = h                                       # ini_file_to_params.f90:303-309
        if not self.CFL_nu:
            digit = self.discretization[3]                       # module_ACM.f90:367-381
            den = {"2": 4.000, "4": 5.333, "6": 6.0444}[digit]
            self.CFL_nu = 0.95 * 2.79 / (den * float(self.dim))
        if (self.penalization or self.use_sponge) and self.n_mask == 0:
            self.n_mask = 6                                      # module_ACM.f90:505
        return self

    @property
    def n_stages(self) -> int:
        return len(self.butcher) - 1

    def is_it_time_to_save_data(self, time: float, iteration: int) -> bool:
        """is_it_time_to_save_data (LIB/IO/save_data.f90:255-288) without the wall-clock clause"""
        import math
        due = False
        if self.write_method == "fixed_freq":
            due = self.write_freq > 0 and iteration % self.write_freq == 0
        elif self.write_method == "fixed_time":
            m = math.fmod(time, self.write_time)
            due = abs(m) < 1.0e-12 or abs(m - self.write_time) < 1.0e-12
        return due and not time + 1.0e-12 < self.write_time_first

    def rkc_coefficients(self):
        """rows s of mu, mu_tilde, nu, gamma_tilde, c for RungeKuttaChebychev: from the parameter file with RKC_custom_scheme = 1
        (ini_file_to_params.f90:629-636), else the tabulated scheme of setup_RKC_coefficients (runge_kutta_chebychev.f90:180 ff) -- which is
        the second-order Chebychev scheme of Sommeijer, Shampine & Verwer (RKC, 1998) with damping eps = 10, evaluated here in closed form;
        it reproduces the reference's tables to 3e-15 (tests/test_oracle_rkc.py compares the sampled rows s = 4, 6, 10, 20).  A Fortran host
        passes its own table rows to wgpu_rkc_step."""
        if not self.RKC_custom_scheme:
            return rkc2_coefficients(self.rkc_s)
        rows = tuple(np.asarray(getattr(self, k), dtype=np.float64) for k in ("RKC_mu", "RKC_mu_tilde", "RKC_nu", "RKC_gamma_tilde", "RKC_c"))
        if any(len(r) != self.rkc_s for r in rows):
            raise ValueError(f"RungeKuttaChebychev: every coefficient row needs s = {self.rkc_s} values")
        return rows

    @classmethod
    def from_ini(cls, path: str) -> "Params":
        ini = IniFile(path)
        p = cls()
        p.dim = ini.integer("Domain", "dim", 2)
        p.domain = tuple((ini.vector("Domain", "domain_size", [1.0, 1.0, 1.0]) + [1.0, 1.0])[:3])
        p.periodic = tuple((ini.vector("Domain", "periodic_BC", [1, 1, 1], int) + [1, 1])[:3])
        p.wavelet = ini.string("Wavelet", "wavelet", "CDF40")
        _, _, g_def, grhs_def = wavelet_ghosts(p.wavelet)
        bs = ini.vector("Blocks", "number_block_nodes", [16], int)
        bs = (bs * 3)[:3] if len(bs) == 1 else (bs + [1])[:3]
        if p.dim == 2:
            bs[2] = 1
        p.Bs = tuple(bs)
        p.g = ini.integer("Blocks", "number_ghost_nodes", g_def)
        p.g_rhs = ini.integer("Blocks", "number_ghost_nodes_rhs", grhs_def)
        p.n_eqn = ini.integer("Blocks", "number_equations", 1)
        p.Jmax = ini.integer("Blocks", "max_treelevel", 5)
        p.Jmin = ini.integer("Blocks", "min_treelevel", 1)
        p.eps = ini.real("Blocks", "eps", 1e-3)
        p.eps_normalized = ini.boolean("Blocks", "eps_normalized", False)
        p.eps_norm = ini.string("Blocks", "eps_norm", "Linfty")
        p.force_maxlevel_dealiasing = ini.boolean("Blocks", "force_maxlevel_dealiasing", False)
        lifted = p.wavelet[4] != "0"
        p.useCoarseExtension = int(ini.boolean("Blocks", "useCoarseExtension", lifted))
        p.useSecurityZone = int(ini.boolean("Blocks", "useSecurityZone", lifted))
        p.adapt_tree = ini.boolean("Blocks", "adapt_tree", False)
        p.refinement_indicator = ini.string("Blocks", "refinement_indicator", "everywhere")
        p.threshold_state_vector_component = tuple(ini.vector("Blocks", "threshold_state_vector_component", [], int))
        p.threshold_mask = ini.boolean("Blocks", "threshold_mask", False)
        p.adapt_inicond = ini.boolean("Blocks", "adapt_inicond", p.adapt_tree)
        p.read_from_files = ini.boolean("Physics", "read_from_files", False)
        p.input_files = tuple(ini.string("Physics", "input_files", "").split())
        p.block_dist = ini.string("Blocks", "block_dist", "sfc_hilbert")
        p.discretization = ini.string("Discretization", "order_discretization", "FD_4th_central")
        p.time_max = ini.real("Time", "time_max", 1.0)
        p.nt = ini.integer("Time", "nt", 99999999)
        p.CFL = ini.real("Time", "CFL", 1.0)
        p.CFL_eta = ini.real("Time", "CFL_eta", 0.99)
        p.CFL_nu = ini.real("Time", "CFL_nu", 0.0)
        p.dt_fixed = ini.real("Time", "dt_fixed", 0.0)
        p.dt_max = ini.real("Time", "dt_max", 0.0)
        p.write_method = ini.string("Time", "write_method", "fixed_freq")
        p.write_time = ini.real("Time", "write_time", 1.0)
        p.write_time_first = ini.real("Time", "write_time_first", 0.0)
        p.write_freq = ini.integer("Time", "write_freq", 25)
        p.tsave_stats = ini.real("Statistics", "tsave_stats", 9999999.9)
        p.nsave_stats = ini.integer("Statistics", "nsave_stats", 99999999)
        p.butcher = ini.matrix("Time", "butcher_tableau", [r[:] for r in BUTCHER_RK4])
        p.time_step_method = ini.string("Time", "time_step_method", "RungeKuttaGeneric")       # ini_file_to_params.f90:592
        p.rkc_s = ini.integer("Time", "s", 4)                                                  # :627
        p.RKC_custom_scheme = ini.boolean("Time", "RKC_custom_scheme", False)                  # :629-636
        if p.RKC_custom_scheme:
            for key in ("RKC_mu", "RKC_mu_tilde", "RKC_nu", "RKC_gamma_tilde", "RKC_c"):
                v = tuple(ini.vector("Time", key, []))
                if len(v) < p.rkc_s:
                    raise ValueError(f"[Time] {key} needs s = {p.rkc_s} values")
                setattr(p, key, v[:p.rkc_s])
        p.M_krylov = ini.integer("Time", "M_krylov", 12)                                       # ini_file_to_params.f90:593-595
        p.krylov_err_threshold = ini.real("Time", "krylov_err_threshold", 1.0e-3)
        p.krylov_subspace_dimension = ini.string("Time", "krylov_subspace_dimension", "fixed")
        p.filter_type = ini.string("Discretization", "filter_type", "no_filter")              # :176-184
        p.filter_only_maxlevel = ini.boolean("Discretization", "filter_only_maxlevel", False)
        p.filter_all_except_maxlevel = ini.boolean("Discretization", "filter_all_except_maxlevel", False)
        if p.filter_type != "no_filter":
            p.filter_freq = ini.integer("Discretization", "filter_freq", -1)
            p.filter_component = tuple(int(x) for x in ini.vector("Discretization", "filter_component", [], int))
        p.c0 = ini.real("ACM-new", "c_0", 10.0)
        p.nu = ini.real("ACM-new", "nu", 1e-1)
        p.gamma_p = ini.real("ACM-new", "gamma_p", 1.0)
        p.u_mean_set = tuple((ini.vector("ACM-new", "u_mean_set", [1.0, 0.0, 0.0]) + [0.0, 0.0])[:3])
        p.skew_symmetry = ini.boolean("ACM-new", "skew_symmetry", False)
        p.inicond = ini.string("ACM-new", "inicond", "meanflow")
        p.penalization = ini.boolean("VPM", "penalization", True)
        p.C_eta = ini.real("VPM", "C_eta", 1.0)
        p.use_sponge = ini.boolean("Sponge", "use_sponge", False)
        p.C_sponge = ini.real("Sponge", "C_sponge", 1.0e-2)
        return p.finalize()

    def to_config(self, max_blocks: int, device: int = 0) -> WgpuConfig:
        self.finalize()
        c = WgpuConfig()
        c.dim = self.dim
        for d in range(3):
            c.Bs[d] = int(self.Bs[d]) if d < self.dim else 1
            c.domain[d] = float(self.domain[d])
            c.periodic[d] = int(self.periodic[d])
            c.u_mean_set[d] = float(self.u_mean_set[d])
        c.g, c.g_rhs, c.n_eqn, c.n_mask = self.g, self.g_rhs, self.n_eqn, self.n_mask
        c.max_blocks, c.Jmax = int(max_blocks), self.Jmax
        c.fd = FD_IDS[self.discretization]
        c.skew_symmetry, c.penalization, c.use_sponge = int(self.skew_symmetry), int(self.penalization), int(self.use_sponge)
        s = self.n_stages
        if s > WGPU_MAX_STAGES:
            raise ValueError("butcher tableau too large")
        c.n_stages = s
        b = np.asarray(self.butcher, dtype=np.float64)
        for i in range(s + 1):
            for j in range(s + 1):
                c.butcher[i * (s + 1) + j] = b[i, j]
        c.write_method_fixed_time = int(self.write_method == "fixed_time")
        c.device = device
        c.c0, c.nu, c.gamma_p, c.C_eta, c.C_sponge = self.c0, self.nu, self.gamma_p, self.C_eta, self.C_sponge
        c.CFL, c.CFL_eta, c.CFL_nu = self.CFL, self.CFL_eta, self.CFL_nu
        c.dt_fixed, c.dt_max, c.time_max = self.dt_fixed, self.dt_max, self.time_max
        c.write_time, c.write_time_first, c.tsave_stats = self.write_time, self.write_time_first, self.tsave_stats
        return c
