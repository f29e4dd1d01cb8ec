"""The *.t time-series files of a WABBIT run (LIB/IO/module_t_files.f90): one row per call, columns in es15.8 separated by ';'
(flush_t_file, :177-201), appended by rank 0.  write_statistics_acm writes the files of STATISTICS_ACM's post_stage
(LIB/EQUATION/ACMnew/statistics_ACM.f90:500-625) that the device statistics cover (wgpu_statistics): umag.t, CFL.t, meanflow.t, div.t, forces.t,
mask_volume.t, penal_power.t, u_residual.t, e_kin.t, enstrophy.t, helicity.t (3-D), dissipation.t (nu > 0).  Host code in a WABBIT build (the
Fortran writer stays as it is and reads the 23 numbers of wgpu_statistics); restated here for the Python time loop."""
from __future__ import annotations

import math
import os
from typing import Mapping, Sequence

TFILE_SEPARATOR = ";"


def format_row(values: Sequence[float]) -> str:
    """one line of a *.t file: '(n-1 (es15.8,";"), es15.8)'"""
    return TFILE_SEPARATOR.join(f"{float(v):15.8E}" for v in values)


def append_t_file(path: str, values: Sequence[float]) -> None:
    with open(path, "a") as f:
        f.write(format_row(values) + "\n")


def write_statistics_acm(stats: Mapping[str, float], time: float, dt: float, p, dx_min: float, directory: str = ".") -> None:
    """the post_stage's append_t_file calls for one statistics_ACM result (`stats`: WabbitGPU.statistics_ACM(..., with_vorticity=True)); `p`:
    Params; dx_min: the smallest lattice spacing of the current grid (the reference reduces it over the blocks)"""
    out = lambda name, row: append_t_file(os.path.join(directory, name), [time] + list(row))   # noqa: E731
    umag = stats["umag"]                                             # max |u|^2
    eig = math.sqrt(umag) + math.sqrt(p.c0 ** 2 + umag)
    out("umag.t", [math.sqrt(umag), p.c0, p.c0 / math.sqrt(umag) if umag > 1.0e-12 else 0.0, eig])
    out("CFL.t", [dt * eig / dx_min, dt * p.nu / dx_min ** 2, dt / p.C_eta])
    out("meanflow.t", [stats["meanflow_x"], stats["meanflow_y"], stats["meanflow_z"]])
    out("div.t", [stats["div_max"], stats["div_min"]])
    if p.penalization or p.use_sponge:
        out("forces.t", [stats["force_x"], stats["force_y"], stats["force_z"]])                  # one geometry: colour 1
        out("mask_volume.t", [stats["mask_volume"], stats["sponge_volume"]])
        out("penal_power.t", [stats["penal_power_solid_input"], stats["penal_power_solid_dissipation"], stats["penal_power_sponge"]])
        out("u_residual.t", [stats["u_residual_x"], stats["u_residual_y"], stats["u_residual_z"]])
    out("e_kin.t", [stats["e_kin"], stats["ACM_energy"]])
    if "enstrophy" in stats:
        out("enstrophy.t", [stats["enstrophy"], stats["max_vort"]])
        if p.dim == 3:
            out("helicity.t", [stats["helicity"]])
        if p.nu > 0.0:
            out("dissipation.t", [stats["dissipation"]])


def statistics_due(iteration: int, time: float, nsave_stats: int, tsave_stats: float) -> bool:
    """main.f90:392 -- every nsave_stats iterations, or when the time is a multiple of tsave_stats (which the time-step control hits exactly)"""
    m = math.fmod(time, tsave_stats)
    return iteration % nsave_stats == 0 or abs(m) < 1.0e-12 or abs(m - tsave_stats) < 1.0e-12
