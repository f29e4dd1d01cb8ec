"""Generate the small golden fixtures under tests/golden/ from the reference checkout.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
The GPU box has no /root/reference, so tests only read the .npz files written here.

taylor_green_{FD2_CDF20,FD4_CDF40,FD6_CDF60}.npz : strided samples (every 3rd interior point) of the reference's own
  regression fields TESTING/acm/taylorGreen/*/{ux,uy,uz,p}_000010000000.h5 (t = 10, written by the reference
  Fortran code), keyed by zero-based block coordinates, plus the iteration count stored in the file.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
from h5lite import read_wabbit  # noqa: E402

REF = "/root/reference/TESTING/acm/taylorGreen"
STRIDE = 3


def main():
    for case in ("taylorGreenEqui_FD2_CDF20", "taylorGreenEqui_FD4_CDF40", "taylorGreenEqui_FD6_CDF60"):
        out = {}
        for tag, key in (("000000000000", "t0"), ("000010000000", "t1")):
            fields = []
            for name in ("ux", "uy", "uz", "p"):
                d = read_wabbit(os.path.join(REF, case, f"{name}_{tag}.h5"))
                Bs = int(d["attrs"]["block-size"][0])
                nb = d["blocks"].shape[0]
                # coords_origin is stored (z,y,x); block coordinate = origin / (Bs*dx)
                ixyz = np.rint(d["origin"][:, ::-1] / (d["spacing"][:, ::-1] * Bs)).astype(np.int32)
                order = np.lexsort((ixyz[:, 0], ixyz[:, 1], ixyz[:, 2]))
                blocks = d["blocks"][order][:, :Bs:STRIDE, :Bs:STRIDE, :Bs:STRIDE]
                fields.append(blocks)
                out[f"{key}_ixyz"] = ixyz[order]
                out[f"{key}_time"] = d["attrs"]["time"]
                out[f"{key}_iteration"] = d["attrs"]["iteration"]
                out["Bs"] = np.array([Bs])
            out[key] = np.stack(fields, axis=1)   # [block, comp, z, y, x] strided
        out["stride"] = np.array([STRIDE])
        path = os.path.join(HERE, case.replace("taylorGreenEqui", "taylor_green") + ".npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path), out["t1"].shape, out["t1_iteration"])




# ------------------------------------------------------------------------------------------------------------------
# TESTING/wavelets: vor_000020000000.h5 --refine-everywhere--> adaptive_CDF40/vor_00100.h5 --coarsen-everywhere-->
# adaptive_CDFxy/vor_00200.h5 (2-D, Bs=32, 4 levels).  For blocks whose whole neighbourhood is on their own level the
# ghost layers are plain copies, so the reference's outputs pin `prediction`/refineBlock (CDF40 refine) and the
# low-pass decomposition filter HD + decimation alignment (CDF44 / CDF42 / CDF22 / CDF62 coarsen) block by block.
# wavelet_blocks.npz holds, for a few such blocks, the ghosted input assembled from the reference's input file and the
# reference's output blocks.
def _grid(d):
    Bs = int(d["attrs"]["block-size"][0])
    L = float(d["attrs"]["domain-size"][0])
    dx = d["spacing"][:, ::-1]            # stored (y,x)
    x0 = d["origin"][:, ::-1]
    level = np.rint(np.log2(L / (Bs * dx[:, 0]))).astype(np.int32)
    ixy = np.rint(x0 / (Bs * dx)).astype(np.int32)
    return Bs, level, ixy


def wavelet_blocks():
    W = "/root/reference/TESTING/wavelets"
    d0 = read_wabbit(os.path.join(W, "vor_000020000000.h5"))
    d1 = read_wabbit(os.path.join(W, "adaptive_CDF40", "vor_00100.h5"))
    Bs, lv0, ix0 = _grid(d0)
    _, lv1, ix1 = _grid(d1)
    look0 = {(int(l), int(i[0]), int(i[1])): k for k, (l, i) in enumerate(zip(lv0, ix0))}
    look1 = {(int(l), int(i[0]), int(i[1])): k for k, (l, i) in enumerate(zip(lv1, ix1))}
    coarsened = {w: read_wabbit(os.path.join(W, f"adaptive_{w}", "vor_00200.h5")) for w in ("CDF22", "CDF42", "CDF44", "CDF62")}
    out = {"Bs": np.array([Bs])}
    G0, G1 = 3, 6
    picked = 0
    for k in range(len(lv0)):
        J, (bx, by) = int(lv0[k]), ix0[k]
        n = 2 ** J
        # the 5x5 neighbourhood must be on the same level: then the 3x3 neighbourhood of every daughter's neighbourhood
        # is uniform as well
        nb = {(dx_, dy_): look0.get((J, (bx + dx_) % n, (by + dy_) % n)) for dx_ in (-2, -1, 0, 1, 2) for dy_ in (-2, -1, 0, 1, 2)}
        if any(v is None for v in nb.values()):
            continue
        # ghosted mother, g = 3 (CDF40 refine input)
        m = np.zeros((Bs + 2 * G0, Bs + 2 * G0))
        big = np.block([[d0["blocks"][nb[(dx_, dy_)]][:Bs, :Bs] for dx_ in (-1, 0, 1)] for dy_ in (-1, 0, 1)])   # [y,x]
        m[:, :] = big[Bs - G0:2 * Bs + G0, Bs - G0:2 * Bs + G0]
        # the reference's four daughters (interiors), digit bit0 -> y, bit1 -> x
        dau = np.zeros((4, Bs, Bs))
        for kd in range(4):
            ox, oy = (kd // 2) % 2, kd % 2
            dau[kd] = d1["blocks"][look1[(J + 1, 2 * bx + ox, 2 * by + oy)]][:Bs, :Bs]
        # composite fine field around the four daughters with a 6-wide ring, from the reference's refined files
        # (the refined data depend on the predictor order X only, and are stored for the unlifted wavelets CDFX0)
        n1 = 2 * n
        for X in (2, 4, 6):
            dX = read_wabbit(os.path.join(W, f"adaptive_CDF{X}0", "vor_00100.h5"))
            _, lvX, ixX = _grid(dX)
            lookX = {(int(l), int(i[0]), int(i[1])): q for q, (l, i) in enumerate(zip(lvX, ixX))}
            fine = np.block([[dX["blocks"][lookX[(J + 1, (2 * bx + ax) % n1, (2 * by + ay) % n1)]][:Bs, :Bs] for ax in (-1, 0, 1, 2)]
                             for ay in (-1, 0, 1, 2)])
            out[f"fine_X{X}_{picked}"] = fine[Bs - G1:3 * Bs + G1, Bs - G1:3 * Bs + G1]
        out[f"mother{picked}"] = m
        out[f"daughters{picked}"] = dau
        for w, dc in coarsened.items():
            _, lvc, ixc = _grid(dc)
            kc = [q for q in range(len(lvc)) if lvc[q] == J and ixc[q][0] == bx and ixc[q][1] == by][0]
            out[f"coarse_{w}_{picked}"] = dc["blocks"][kc][:Bs, :Bs]
        picked += 1
        if picked == 3:
            break
    out["n"] = np.array([picked])
    path = os.path.join(HERE, "wavelet_blocks.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), picked)



# ------------------------------------------------------------------------------------------------------------------
# TESTING/acm/3vortices: 2-D ACM, Bs=32, equidistant level 3 (64 blocks), skew-symmetric, restart from the stored
# {ux,uy,p}_000010000000.h5 (t = 10) and run to t = 20.  three_vortices_t10.npz holds the full restart fields (interior
# points), three_vortices_FDx_CDFy0.npz strided samples of the reference's t = 20 output and its iteration counter.
def three_vortices():
    R = "/root/reference/TESTING/acm/3vortices"
    out = {}
    fields = []
    for name in ("ux", "uy", "p"):
        d = read_wabbit(os.path.join(R, f"{name}_000010000000.h5"))
        Bs = int(d["attrs"]["block-size"][0])
        ixy = np.rint(d["origin"][:, ::-1] / (d["spacing"][:, ::-1] * Bs)).astype(np.int32)
        order = np.lexsort((ixy[:, 0], ixy[:, 1]))
        fields.append(d["blocks"][order][:, :Bs, :Bs])
        out["ixy"] = ixy[order]
        out["iteration"] = d["attrs"]["iteration"]
        out["time"] = d["attrs"]["time"]
        out["level"] = d["level"][order]
    out["u"] = np.stack(fields, axis=1)
    path = os.path.join(HERE, "three_vortices_t10.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), out["u"].shape)
    for case in ("3vorticesEquiFD2_CDF20", "3vorticesEquiFD4_CDF40", "3vorticesEquiFD6_CDF60"):
        o = {}
        fields = []
        for name in ("ux", "uy", "p"):
            d = read_wabbit(os.path.join(R, case, f"{name}_000020000000.h5"))
            Bs = int(d["attrs"]["block-size"][0])
            ixy = np.rint(d["origin"][:, ::-1] / (d["spacing"][:, ::-1] * Bs)).astype(np.int32)
            order = np.lexsort((ixy[:, 0], ixy[:, 1]))
            fields.append(d["blocks"][order][:, :Bs:STRIDE, :Bs:STRIDE])
            o["ixy"] = ixy[order]
            o["iteration"] = d["attrs"]["iteration"]
            o["time"] = d["attrs"]["time"]
        o["u"] = np.stack(fields, axis=1)
        o["stride"] = np.array([STRIDE])
        path = os.path.join(HERE, case.replace("3vorticesEqui", "three_vortices_") + ".npz")
        np.savez_compressed(path, **o)
        print(path, os.path.getsize(path), o["u"].shape, o["iteration"])


# ------------------------------------------------------------------------------------------------------------------
# TESTING/acm/3vortices/3vorticesAdaptFD{2,4,6}_CDF{20,22,40,42,60,62} (the parameter files differ in wavelet and order_discretization only):
# the same restart, adaptive (eps = 1e-3, Jmin 1, Jmax 4, refinement
# indicator "significant", coarse extension and security zone on, c_0 = 5): the stored grid after adapt_inicond (t = 10) and
# after 2281 adaptive steps (t = 15).  Per file: block levels, zero-based block coordinates, the stored refinement status of
# every block (0 significant / 9 REF_UNSIGNIFICANT_STAY), iteration, time; the fields in full at t = 10 (25 blocks) and as strided
# samples at t = 15.
def three_vortices_adaptive(cases=("3vorticesAdaptFD4_CDF40", "3vorticesAdaptFD4_CDF42", "3vorticesAdaptFD2_CDF20", "3vorticesAdaptFD2_CDF22",
                                   "3vorticesAdaptFD6_CDF60", "3vorticesAdaptFD6_CDF62")):
    R = "/root/reference/TESTING/acm/3vortices"
    for case in cases:
        o = {}
        for tag, key, stride in (("000010000000", "t10", 1), ("000015000000", "t15", 2)):
            fields = []
            for name in ("ux", "uy", "p"):
                d = read_wabbit(os.path.join(R, case, f"{name}_{tag}.h5"))
                Bs = int(d["attrs"]["block-size"][0])
                ixy = np.rint(d["origin"][:, ::-1] / (d["spacing"][:, ::-1] * Bs)).astype(np.int32)
                lvl = d["level"].ravel().astype(np.int32)
                order = np.lexsort((ixy[:, 1], ixy[:, 0], lvl))
                fields.append(d["blocks"][order][:, :Bs:stride, :Bs:stride])
                o[f"{key}_ixy"] = ixy[order]
                o[f"{key}_level"] = lvl[order]
                o[f"{key}_status"] = d["refinement_status"].ravel().astype(np.int32)[order]
                o[f"{key}_iteration"] = d["attrs"]["iteration"]
                o[f"{key}_time"] = d["attrs"]["time"]
            o[f"{key}_u"] = np.stack(fields, axis=1)
            o[f"{key}_stride"] = np.array([stride])
            # the derived fields the reference saved with the state (field_names vor, div: compute_vorticity / divergence with the case's
            # stencils on the synchronised state, PREPARE_SAVE_DATA_ACM): every 2nd point; at t = 15 together with the full velocity of the
            # two cases FD4_CDF42 / FD6_CDF62 (the strided samples above cannot be differentiated)
            for name in ("vor", "div"):
                d = read_wabbit(os.path.join(R, case, f"{name}_{tag}.h5"))
                o[f"{key}_{name}"] = d["blocks"][order][:, :Bs:2, :Bs:2]
            if key == "t15" and case in ("3vorticesAdaptFD4_CDF42", "3vorticesAdaptFD6_CDF62"):
                full = []
                for name in ("ux", "uy"):
                    d = read_wabbit(os.path.join(R, case, f"{name}_{tag}.h5"))
                    full.append(d["blocks"][order][:, :Bs, :Bs])
                o["t15_velocity"] = np.stack(full, axis=1)
        path = os.path.join(HERE, case.replace("3vorticesAdapt", "three_vortices_adapt_") + ".npz")
        np.savez_compressed(path, **o)
        print(path, os.path.getsize(path), o["t10_u"].shape, o["t15_u"].shape, o["t15_iteration"])


# ------------------------------------------------------------------------------------------------------------------
# TESTING/acm/acm_CDF44 (acm_cyl.ini): 2-D flow past a cylinder, Bs = 26, CDF44, Jmax = 6, adaptive every step, penalization (C_eta =
# 1.34e-3, cosine-smoothed circle of radius 0.5 at (10, 10)), p-norm sponge, threshold_mask, force_maxlevel_dealiasing; the stored grids at
# t = 0 (after the adaptive initial condition), 0.05 (iteration 40) and 0.1 (iteration 82).  Per time: block levels, zero-based block
# coordinates, refinement statuses, iteration, time, strided samples of ux, uy, p and of the mask function.
def cylinder_adaptive():
    """acm_CDF44 as described above; acm_CDF40: the same with the unlifted CDF40 (no coarse extension, no security zone);
    acm_norm_CDF44: Jmax = 5, threshold_state_vector_component = 2 2 1 (joint norm of the velocity), one file at t = 0.2;
    acm_significant_CDF44: Jmax = 5, refinement_indicator = significant, one file at t = 0.2"""
    for case, tags in (("acm_CDF44", (("000000000000", "t0"), ("000000050000", "t1"), ("000000100000", "t2"))),
                       ("acm_CDF40", (("000000000000", "t0"), ("000000050000", "t1"), ("000000100000", "t2"))),
                       ("acm_norm_CDF44", (("000000000000", "t0"), ("000000200000", "t2"))),
                       ("acm_significant_CDF44", (("000000000000", "t0"), ("000000200000", "t2")))):
        R = "/root/reference/TESTING/acm/" + case
        o = {}
        for tag, key in tags:
            fields = []
            for name in ("ux", "uy", "p", "mask"):
                d = read_wabbit(os.path.join(R, f"{name}_{tag}.h5"))
                Bs = int(d["attrs"]["block-size"][0])
                ixy = np.rint(d["origin"][:, ::-1] / (d["spacing"][:, ::-1] * Bs)).astype(np.int32)
                lvl = d["level"].ravel().astype(np.int32)
                order = np.lexsort((ixy[:, 1], ixy[:, 0], lvl))
                fields.append(d["blocks"][order][:, :Bs:2, :Bs:2])
                o[f"{key}_ixy"] = ixy[order]
                o[f"{key}_level"] = lvl[order]
                o[f"{key}_status"] = d["refinement_status"].ravel().astype(np.int32)[order]
                o[f"{key}_iteration"] = d["attrs"]["iteration"]
                o[f"{key}_time"] = d["attrs"]["time"]
            o[f"{key}_u"] = np.stack(fields[:3], axis=1)
            o[f"{key}_mask"] = fields[3]
            o[f"{key}_stride"] = np.array([2])
        path = os.path.join(HERE, case.replace("acm_", "cylinder_adapt_") + ".npz")
        np.savez_compressed(path, **o)
        print(path, os.path.getsize(path), [o[f"{k}_u"].shape for _, k in tags], o["t2_iteration"])


if __name__ == "__main__":
    main()
    wavelet_blocks()
    three_vortices()
    three_vortices_adaptive()
    cylinder_adaptive()


# ------------------------------------------------------------------------------------------------------------------
# Runge-Kutta-Chebychev coefficient tables (LIB/TIME/runge_kutta_chebychev.f90: setup_RKC_coefficients): the rows s = 4, 6, 10, 20 of
# mu, mu_tilde, nu, gamma_tilde, c as the reference's source lists them (damping eps = 10) -> rkc_coefficients.npz
def rkc_tables(stages=(4, 6, 10, 20)):
    import re
    src = open("/root/reference/LIB/TIME/runge_kutta_chebychev.f90").read()
    out = {}
    for s in stages:
        m = re.search(r"\n\s*s=%d\n(.*?)(?=\n\s*! -{10,}|\nend subroutine)" % s, src, flags=re.S)
        body = m.group(1)
        for name in ("mu", "mu_tilde", "nu", "gamma_tilde", "c"):
            mm = re.search(r"\b%s\(s,1:%d\)=\(/(.*?)/\)" % (name, s), body, flags=re.S)
            vals = [float(v.replace("_rk", "")) for v in re.findall(r"[-+]?\d\.\d+e[-+]\d+_rk", mm.group(1))]
            assert len(vals) == s, (s, name, len(vals))
            out[f"s{s}_{name}"] = np.array(vals)
    path = os.path.join(HERE, "rkc_coefficients.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), sorted(out)[:5])


if __name__ == "__main__":
    rkc_tables()


# ------------------------------------------------------------------------------------------------------------------
# TESTING/wavelets, whole files (runtests.py, group "adaptive"): vor_000020000000.h5 (2-D, Bs = 32, 112 blocks on levels 2-5, one
# component) --wabbit-post --refine-everywhere--> adaptive_CDFX0/vor_00100.h5 (448 blocks) --wabbit-post --coarsen-everywhere-->
# adaptive_CDFXY/vor_00200.h5 (112 blocks) for the seven wavelets CDF20 / 22 / 40 / 42 / 44 / 60 / 62 (sparse_to_dense.f90:183-235:
# sync_ghosts_tree, refine_tree("everywhere") resp. adapt_tree("everywhere"), useCoarseExtension = useSecurityZone = isLiftedWavelet).
# wavelet_files.npz: the input in full; of every output the block list, strided samples (every 4th / 2nd point) and, for the refined
# files, the SHA-256 of the float64 interiors in (level, ix, iy) order -- the oracle reproduces those bit for bit.
def wavelet_files():
    import hashlib
    W = "/root/reference/TESTING/wavelets"

    def load(path):
        d = read_wabbit(path)
        Bs, level, ixy = _grid(d)
        order = np.lexsort((ixy[:, 1], ixy[:, 0], level))
        return Bs, level[order], ixy[order], np.ascontiguousarray(d["blocks"][order][:, :Bs, :Bs])
    Bs, lv, ix, blk = load(os.path.join(W, "vor_000020000000.h5"))
    out = {"Bs": np.array([Bs]), "in_level": lv, "in_ixy": ix, "in_blocks": blk}
    for X in (2, 4, 6):
        _, lv1, ix1, b1 = load(os.path.join(W, f"adaptive_CDF{X}0", "vor_00100.h5"))
        out[f"refined_X{X}_level"], out[f"refined_X{X}_ixy"] = lv1, ix1
        out[f"refined_X{X}_sample"] = b1[:, ::4, ::4]
        out[f"refined_X{X}_sha256"] = np.frombuffer(hashlib.sha256(b1.astype("<f8").tobytes()).digest(), dtype=np.uint8)
    for w in ("CDF20", "CDF22", "CDF40", "CDF42", "CDF44", "CDF60", "CDF62"):
        _, lv2, ix2, b2 = load(os.path.join(W, f"adaptive_{w}", "vor_00200.h5"))
        out[f"coarsened_{w}_level"], out[f"coarsened_{w}_ixy"] = lv2, ix2
        out[f"coarsened_{w}_sample"] = b2[:, ::2, ::2]
    path = os.path.join(HERE, "wavelet_files.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    wavelet_files()


# ------------------------------------------------------------------------------------------------------------------
# TESTING/acm/bumblebeeFlowEquiFD4_CDF40 (3-D, Bs = 26, level 1 = 8 blocks, periodic, FD_4th_central): the reference saved |vorticity|
# (field "vorabs": compute_vorticity_abs, LIB/OPERATORS/compute_vorticity.f90:70-123) next to ux, uy, uz.  The flow itself needs the insect
# module, but the derived field is a golden vector for the vorticity operator: vorabs_3d.npz holds, at t = 2, two blocks' velocity with a
# two-point halo taken from their (periodic) neighbours and the reference's vorabs of those blocks.
def vorabs_3d():
    Bd = "/root/reference/TESTING/acm/bumblebeeFlowEquiFD4_CDF40"
    tag, H = "000002000000", 2

    def load(name):
        d = read_wabbit(os.path.join(Bd, f"{name}_{tag}.h5"))
        Bs = int(d["attrs"]["block-size"][0])
        ixyz = np.rint(d["origin"][:, ::-1] / (d["spacing"][:, ::-1] * Bs)).astype(np.int64)
        return Bs, {tuple(int(v) for v in x): d["blocks"][k][:Bs, :Bs, :Bs] for k, x in enumerate(ixyz)}, float(d["spacing"][0, 0])
    Bs, ux, dx = load("ux")
    _, uy, _ = load("uy")
    _, uz, _ = load("uz")
    _, va, _ = load("vorabs")
    out = {"Bs": np.array([Bs]), "H": np.array([H]), "dx": np.array([dx])}
    for j, blk in enumerate(((0, 0, 0), (1, 0, 1))):
        comp = []
        for f in (ux, uy, uz):
            big = np.block([[[f[((blk[0] + ax) % 2, (blk[1] + ay) % 2, (blk[2] + az) % 2)] for ax in (-1, 0, 1)] for ay in (-1, 0, 1)]
                            for az in (-1, 0, 1)])                      # [z, y, x]
            comp.append(big[Bs - H:2 * Bs + H, Bs - H:2 * Bs + H, Bs - H:2 * Bs + H])
        out[f"u{j}"] = np.stack(comp)
        out[f"vorabs{j}"] = va[blk]
    path = os.path.join(HERE, "vorabs_3d.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), float(out["vorabs0"].max()))


if __name__ == "__main__":
    vorabs_3d()


# ------------------------------------------------------------------------------------------------------------------
# TESTING/acm/3vortices/3vorticesEqui*/log.original.txt: the reference's own log of the equidistant runs, one line per time step
# ("RUN: it= 3055 time= 10.003239697 ... dt= 3.2E-03"): the time after every one of the 3073 steps to nine decimals -- a golden vector
# for calculate_time_step along the whole run.  three_vortices_log_times.npz: iteration and time per step and case.
def three_vortices_logs():
    import re
    R = "/root/reference/TESTING/acm/3vortices"
    out = {}
    for case in ("FD2_CDF20", "FD4_CDF40", "FD6_CDF60"):
        it, tm = [], []
        for line in open(os.path.join(R, f"3vorticesEqui{case}", "log.original.txt"), errors="replace"):
            m = re.match(r"RUN: it=\s*(\d+) time=\s*([0-9.]+) ", line)
            if m:
                it.append(int(m.group(1)))
                tm.append(float(m.group(2)))
        out[f"{case}_iteration"], out[f"{case}_time"] = np.array(it, dtype=np.int64), np.array(tm)
        print(case, len(it), it[0], it[-1], tm[-1])
    path = os.path.join(HERE, "three_vortices_log_times.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    three_vortices_logs()
