"""Shared helpers of the parity tests: build the same case for the product (wabbit_b200) and for the oracle."""
import numpy as np

import oracle as O
from wabbit_b200 import Forest, Params


def orc_params(p: Params) -> O.Params:
    p.finalize()
    return O.Params(dim=p.dim, Bs=tuple(p.Bs), g=p.g, g_rhs=p.g_rhs, n_eqn=p.n_eqn, domain=tuple(p.domain), Jmax=p.Jmax,
                    discretization=p.discretization, skew=p.skew_symmetry, penalization=p.penalization,
                    use_sponge=p.use_sponge, c0=p.c0, nu=p.nu, gamma_p=p.gamma_p, C_eta=p.C_eta, C_sponge=p.C_sponge,
                    u_mean_set=tuple(p.u_mean_set), CFL=p.CFL, CFL_eta=p.CFL_eta, CFL_nu=p.CFL_nu, dt_fixed=p.dt_fixed,
                    dt_max=p.dt_max, time_max=p.time_max, write_method=p.write_method, write_time=p.write_time,
                    write_time_first=p.write_time_first, tsave_stats=p.tsave_stats, butcher=np.asarray(p.butcher, dtype=np.float64))


def orc_grid(forest: Forest, rank: int = 0) -> O.Grid:
    """Oracle grid whose block k is the forest's k-th active block (hvy id = k+1 on a single rank)."""
    hvy, lvl, ixyz, _ = forest.active(rank)
    assert (hvy == np.arange(1, len(hvy) + 1)).all()
    return O.Grid(level=lvl.astype(np.int64), ixyz=ixyz.astype(np.int64), dim=forest.dim)


def tg_params(Bs=16, J=2, wavelet_g=3, discretization="FD_4th_central", skew=True, **kw) -> Params:
    p = Params(dim=3, domain=(6.283185307179586,) * 3, Bs=(Bs, Bs, Bs), g=wavelet_g, g_rhs=2, n_eqn=4, Jmax=J,
               discretization=discretization, skew_symmetry=skew, c0=10.0, nu=3.125e-3, gamma_p=0.0, CFL=1.0,
               u_mean_set=(0.0, 0.0, 0.0), time_max=1.0e9, **kw)
    return p.finalize()


def relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def graded_blocks(dim, J0, Jmax, seed, frac=0.3, rounds=None):
    """Random graded leaf grid: start equidistant on level J0, refine a random fraction of the leaves `rounds` times
    (default Jmax-J0), restoring gradedness over all 3^dim-1 neighbour directions after every round (what
    ensureGradedness_tree guarantees, LIB/MESH/ensureGradedness_tree.f90:13).  Returns (level[n], ixyz[n,3])."""
    rng = np.random.default_rng(seed)
    n0 = 2 ** J0
    leaves = {(J0, x, y, z) for z in range(n0 if dim == 3 else 1) for y in range(n0) for x in range(n0)}
    dirs = [(dx, dy, dz) for dz in ((-1, 0, 1) if dim == 3 else (0,)) for dy in (-1, 0, 1) for dx in (-1, 0, 1) if (dx, dy, dz) != (0, 0, 0)]

    def refine(b):
        L, x, y, z = b
        leaves.remove(b)
        for c in range(2 ** dim):
            leaves.add((L + 1, 2 * x + (c & 1), 2 * y + ((c >> 1) & 1), 2 * z + ((c >> 2) & 1) if dim == 3 else 0))

    def owner(L, x, y, z):
        for l in range(L, -1, -1):
            s = L - l
            k = (l, x >> s, y >> s, z >> s)
            if k in leaves:
                return k
        return None

    for _ in range((Jmax - J0) if rounds is None else rounds):
        cand = sorted(b for b in leaves if b[0] < Jmax)
        pick = [b for b in cand if rng.random() < frac]
        for b in pick:
            if b in leaves:
                refine(b)
        changed = True
        while changed:
            changed = False
            for b in sorted(leaves):
                if b not in leaves:
                    continue
                L, x, y, z = b
                n = 2 ** L
                for d in dirs:
                    o = owner(L, (x + d[0]) % n, (y + d[1]) % n, (z + d[2]) % n if dim == 3 else 0)
                    if o is not None and o[0] < L - 1:
                        refine(o)
                        changed = True
    lv = np.array([b[0] for b in sorted(leaves)], dtype=np.int32)
    ix = np.array([b[1:] for b in sorted(leaves)], dtype=np.int32)
    return lv, ix
