"""The reference's 2-D adaptive regression cases TESTING/acm/3vortices/3vorticesAdaptFD{2,4,6}_CDF{20,22,40,42,60,62} (PARAMS_3vortices.ini),
shared by the oracle pin (test_oracle_adaptive.py: all six) and the GPU run (test_gpu_adaptive2d.py: FD4_CDF40 / FD4_CDF42).

  restart from {ux,uy,p}_000010000000.h5 (64 blocks of 32^2 on level 3, t = 10, iteration 3054)
  adapt_inicond = 1 -> one adapt_tree;  then main.f90's loop (sync -> refine_tree("significant") -> RK4 -> adapt_tree) to t = 15
  eps = 1e-3 (Linfty, normalised, all three components), Jmin = 1, Jmax = 4, useCoarseExtension = useSecurityZone = 1 (also for the
  unlifted CDF40), FD_4th_central, skew-symmetric, c_0 = 5, nu = 5e-5, gamma_p = 1, CFL = 1, write_time = 10.
"""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(__file__), "golden")
WAVELET_G = {"CDF40": 3, "CDF42": 4}       # setup_wavelet: g = widest filter (module_wavelets.f90:1330-1339)
BS = 32
INI = dict(dim=2, Bs=(BS, BS, 1), g_rhs=2, n_eqn=3, domain=(6.283185307179586,) * 3, Jmax=4, discretization="FD_4th_central",
           c0=5.0, nu=5.0e-5, gamma_p=1.0, CFL=1.0, time_max=15.0, write_method="fixed_time", write_time=10.0, u_mean_set=(0.0, 0.0, 0.0))
EPS, JMIN = 1.0e-3, 1


def restart_fields():
    """levels [64], block coordinates [64, 3], interiors [64, 3, 32, 32], time, iteration of the stored t = 10 restart"""
    inp = np.load(os.path.join(GOLD, "three_vortices_t10.npz"))
    ixyz = np.concatenate([inp["ixy"], np.zeros((len(inp["ixy"]), 1), np.int32)], axis=1).astype(np.int64)
    return inp["level"].astype(np.int64), ixyz, inp["u"], float(inp["time"][0]), int(inp["iteration"][0])


# every adaptive 3vortices case of the reference: directory 3vorticesAdapt<key>, (wavelet, order_discretization); the parameter files differ in
# these two entries only.  g = the wavelet's ghost nodes (setup_wavelet), g_rhs = the stencil half width.
CASES = {"FD4_CDF40": ("CDF40", "FD_4th_central"), "FD4_CDF42": ("CDF42", "FD_4th_central"), "FD2_CDF20": ("CDF20", "FD_2nd_central"),
         "FD2_CDF22": ("CDF22", "FD_2nd_central"), "FD6_CDF60": ("CDF60", "FD_6th_central"), "FD6_CDF62": ("CDF62", "FD_6th_central")}
FD_HALF_WIDTH = {"FD_2nd_central": 1, "FD_4th_central": 2, "FD_6th_central": 3}
CASE_G = {"FD4_CDF40": 3, "FD4_CDF42": 4, "FD2_CDF20": 1, "FD2_CDF22": 2, "FD6_CDF60": 5, "FD6_CDF62": 6}


def case_ini(case: str) -> dict:
    ini = dict(INI)
    ini["discretization"] = CASES[case][1]
    ini["g_rhs"] = FD_HALF_WIDTH[CASES[case][1]]
    return ini


def gold(name: str):
    """the stored files of a case ("FD2_CDF22", ...; a bare wavelet name means the FD4 case)"""
    case = name if name in CASES else "FD4_" + name
    return np.load(os.path.join(GOLD, f"three_vortices_adapt_{case}.npz"))


def compare(gd, key: str, level, ixyz, status, interiors, iteration=None, time=None):
    """grid (level, ixyz[:, :2]) + refinement status identical to the stored file `key` ("t10" / "t15"); returns max |field difference|.
    interiors: [nb, 3, 32, 32]"""
    mine = {(int(l), int(x[0]), int(x[1])): b for b, (l, x) in enumerate(zip(level, ixyz))}
    ref = [(int(l), int(x[0]), int(x[1])) for l, x in zip(gd[f"{key}_level"], gd[f"{key}_ixy"])]
    assert set(mine) == set(ref), (sorted(set(mine) - set(ref))[:5], sorted(set(ref) - set(mine))[:5])
    if iteration is not None:
        assert iteration == int(gd[f"{key}_iteration"][0])
    if time is not None:
        assert time == float(gd[f"{key}_time"][0])
    s = int(gd[f"{key}_stride"][0])
    err = 0.0
    for j, k in enumerate(ref):
        b = mine[k]
        assert int(status[b]) == int(gd[f"{key}_status"][j]), (k, int(status[b]), int(gd[f"{key}_status"][j]))
        err = max(err, float(np.abs(interiors[b][:, ::s, ::s] - gd[f"{key}_u"][j]).max()))
    return err
