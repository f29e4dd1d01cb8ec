"""The reference's 2-D adaptive regression case TESTING/acm/acm_CDF44 (acm_cyl.ini; BASELINE config "2D ACM test case from TESTING/acm",
SURVEY 8d config 1), shared by the oracle pin (test_oracle_cylinder.py) and the GPU run (test_gpu_cylinder2d.py).

  domain 20 x 20 periodic, Bs = 26, CDF44 (g = 6), FD_4th_central, Jmax = 6, Jmin = 1, eps = 1e-3 (Linfty, normalised), threshold_mask = 1,
  force_maxlevel_dealiasing = 1, refinement "everywhere", ACM c_0 = 12.5, nu = 0, gamma_p = 0, u_mean_set = (0, -1), inicond meanflow,
  penalization (cylinder R = 0.5 at (10, 10), cosine smoothing 1.5 dx_min, C_eta = 1.34e-3), sponge p-norm (p = 8, L = 2, C = 8e-3),
  CFL = 1.5, CFL_eta = 0.99, RK4, write_time 0.05, time_max 0.1; adaptive initial condition, then adaptive every step.
"""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(__file__), "golden")
BS, G = 26, 6
INI = dict(dim=2, Bs=(BS, BS, 1), g=G, g_rhs=2, n_eqn=3, domain=(20.0, 20.0, 20.0), Jmax=6, discretization="FD_4th_central", penalization=True,
           use_sponge=True, c0=12.5, nu=0.0, gamma_p=0.0, C_eta=1.34e-3, C_sponge=8.0e-3, u_mean_set=(0.0, -1.0, 0.0), CFL=1.5, CFL_eta=0.99,
           time_max=0.1, write_method="fixed_time", write_time=0.05)
EPS, JMIN = 1.0e-3, 1

# the four stored variants of the case (TESTING/acm/<dir>/acm_cyl.ini); `files`: key -> time of the stored grids
CASES = {
    "CDF44": dict(wavelet="CDF44", g=6, Jmax=6, thresh_comp=None, indicator="everywhere", time_max=0.1, write_time=0.05,
                  files={"t0": 0.0, "t1": 0.05, "t2": 0.1}, nb_rhs_max=640),
    # unlifted: useCoarseExtension = useSecurityZone = isLiftedWavelet = 0 -> adapt_tree keeps the original values
    "CDF40": dict(wavelet="CDF40", g=3, Jmax=6, thresh_comp=None, indicator="everywhere", time_max=0.1, write_time=0.05,
                  files={"t0": 0.0, "t1": 0.05, "t2": 0.1}, nb_rhs_max=604),
    # threshold_state_vector_component = 2 2 1: ux and uy thresholded together with their joint norm
    "norm_CDF44": dict(wavelet="CDF44", g=6, Jmax=5, thresh_comp=(2, 2, 1), indicator="everywhere", time_max=0.2, write_time=0.2,
                       files={"t0": 0.0, "t2": 0.2}, nb_rhs_max=None),
    "significant_CDF44": dict(wavelet="CDF44", g=6, Jmax=5, thresh_comp=None, indicator="significant", time_max=0.2, write_time=0.2,
                              files={"t0": 0.0, "t2": 0.2}, nb_rhs_max=None),
}


def ini(case: str) -> dict:
    c = CASES[case]
    d = dict(INI)
    d.update(g=c["g"], Jmax=c["Jmax"], time_max=c["time_max"], write_time=c["write_time"])
    return d


def gold(case: str = "CDF44"):
    return np.load(os.path.join(GOLD, f"cylinder_adapt_{case}.npz"))


def compare(gd, key: str, level, ixyz, status, interiors, iteration, time, mask_chi=None, check_status: bool = True):
    """grid, refinement statuses, iteration counter and time identical to the stored file; returns max |field difference| (and checks the
    mask function if given: [nb, Bs, Bs])"""
    mine = {(int(l), int(x[0]), int(x[1])): b for b, (l, x) in enumerate(zip(level, ixyz))}
    ref = [(int(l), int(x[0]), int(x[1])) for l, x in zip(gd[f"{key}_level"], gd[f"{key}_ixy"])]
    assert set(mine) == set(ref), (len(mine), len(ref), sorted(set(mine) - set(ref))[:5], sorted(set(ref) - set(mine))[:5])
    assert iteration == int(gd[f"{key}_iteration"][0])
    assert abs(time - float(gd[f"{key}_time"][0])) <= 1e-15
    s = int(gd[f"{key}_stride"][0])
    err = 0.0
    for j, k in enumerate(ref):
        b = mine[k]
        if check_status:
            assert int(status[b]) == int(gd[f"{key}_status"][j]), (k, int(status[b]), int(gd[f"{key}_status"][j]))
        err = max(err, float(np.abs(interiors[b][:, ::s, ::s] - gd[f"{key}_u"][j]).max()))
        if mask_chi is not None:
            assert np.array_equal(mask_chi[b][::s, ::s], gd[f"{key}_mask"][j]), k
    return err
