"""CPU-side checks of the host logic: the C-ABI libraries load and export every declared symbol, the .ini reader,
and the forest tables (hvy_neighbor conventions) against an independent NumPy restatement."""
import ctypes as C
import os
import sys
import re

import numpy as np
import pytest

import oracle as O
from wabbit_b200 import Forest, Params, _build, _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(w(?:gpu|host)_[a-z_0-9]+)\s*\(", src)))


def test_gpu_library_exports_every_declared_symbol():
    assert os.path.exists(_build.GPU_LIB), "libwabbit_gpu.so not built (python -m wabbit_b200._build)"
    lib = C.CDLL(_build.GPU_LIB)   # loading needs no GPU
    names = _declared("wabbit_gpu.h")
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_native.GPU_SYMBOLS), set(names) ^ set(_native.GPU_SYMBOLS)


def test_fortran_bridge_matches_the_header():
    """fortran/module_gpu_bridge.f90 (the bind(C) module a WABBIT maintainer adds; no Fortran compiler here) is exactly what
    fortran/gen_bridge.py generates from include/wabbit_gpu.h: one interface per exported function, the wgpu_config mirror in the header's
    field order and types -- which is also the ctypes mirror's layout, byte for byte."""
    sys.path.insert(0, os.path.join(ROOT, "fortran"))
    import gen_bridge
    text = gen_bridge.generate()
    assert open(gen_bridge.OUT).read() == text, "regenerate: python fortran/gen_bridge.py"
    names = _declared("wabbit_gpu.h")
    for n in names:
        assert f'bind(C, name="{n}")' in text, n
    fields = gen_bridge.parse_config(gen_bridge.strip_comments(open(gen_bridge.HEADER).read()))
    cty = {"int32_t": C.c_int32, "double": C.c_double}
    mirror = _native.WgpuConfig._fields_
    assert [f[0] for f in fields] == [m[0] for m in mirror]
    size = 0
    for (name, ctype, count), (mname, mtype) in zip(fields, mirror):
        n = 1 if count is None else eval(count.replace("WGPU_MAX_STAGES", str(_native.WGPU_MAX_STAGES)))
        want = cty[ctype] if count is None else cty[ctype] * n
        assert C.sizeof(mtype) == C.sizeof(want) and (mtype is want or mtype._type_ is cty[ctype]), name
        size += C.sizeof(want)
    assert C.sizeof(_native.WgpuConfig) == size        # no padding: 20 x int32 before the first double


def test_host_library_exports_every_declared_symbol():
    lib = _native.host_lib()
    names = _declared("wabbit_host.h")
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_native.HOST_SYMBOLS)


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from wabbit_b200 import WabbitAbort, WabbitGPU
    with pytest.raises(WabbitAbort) as e:
        WabbitGPU(Params(), max_blocks=8)
    assert e.value.code == 1004 and "no CPU fallback" in str(e.value)


def test_ini_reader(tmp_path):
    ini = tmp_path / "PARAMS.ini"
    ini.write_text("""
; comment
[Domain]
dim=3;
domain_size=6.283185307179586 6.283185307179586 6.283185307179586;
periodic_BC=1 1 1;
[Wavelet]
wavelet=CDF44; some comment
[Blocks]
number_block_nodes=16;
number_ghost_nodes=;
number_ghost_nodes_rhs=1;
number_equations=4;
max_treelevel=4;
[Time]
CFL=1.0;
CFL_nu=;
write_method=fixed_time;
write_time=10.0;
butcher_tableau=(/ 0.0 0.0 0.0
0.5 0.5 0.0
0.0 0.0 1.0 /)
[ACM-new]
c_0=10;
nu=3.125000e-03;
gamma_p=0;
skew_symmetry=1;
[Discretization]
order_discretization=FD_4th_central;
[VPM]
penalization=0;
""")
    p = Params.from_ini(str(ini))
    assert p.dim == 3 and p.Bs == (16, 16, 16) and p.wavelet == "CDF44"
    assert p.g == 6            # CDF44: X-1 + Y-1            (ini_file_to_params.f90:467)
    assert p.g_rhs == 2        # raised from 1 to the FD4 half width (ini_file_to_params.f90:303-309)
    assert p.skew_symmetry and not p.penalization and p.n_mask == 0
    assert abs(p.CFL_nu - 0.95 * 2.79 / (5.333 * 3)) < 1e-15
    assert p.n_stages == 2 and p.butcher[1] == [0.5, 0.5, 0.0]
    cfg = p.to_config(64)
    assert cfg.n_stages == 2 and cfg.butcher[3] == 0.5 and cfg.write_method_fixed_time == 1


def test_odd_block_size_rejected():
    with pytest.raises(ValueError):
        Params(Bs=(17, 17, 17)).finalize()


# slot -> direction, from the lists in get_indices_of_modify_patch (LIB/TREE/neighborhood.f90:75-92)
def _dir_of_slot(s):
    d = [0, 0, 0]
    if s in (1, 2, 3, 4, 25, 26, 29, 30, 33, 34, 37, 38, 49, 51, 53, 55): d[0] = -1
    if s in (5, 6, 7, 8, 27, 28, 31, 32, 35, 36, 39, 40, 50, 52, 54, 56): d[0] = +1
    if s in (9, 10, 11, 12, 25, 26, 27, 28, 41, 42, 45, 46, 49, 50, 53, 54): d[1] = -1
    if s in (13, 14, 15, 16, 29, 30, 31, 32, 43, 44, 47, 48, 51, 52, 55, 56): d[1] = +1
    if s in (17, 18, 19, 20, 33, 34, 35, 36, 41, 42, 43, 44, 49, 50, 51, 52): d[2] = -1
    if s in (21, 22, 23, 24, 37, 38, 39, 40, 45, 46, 47, 48, 53, 54, 55, 56): d[2] = +1
    return d


@pytest.mark.parametrize("dim,J,sfc", [(3, 2, "sfc_hilbert"), (3, 3, "sfc_z"), (2, 3, "sfc_hilbert")])
def test_uniform_forest_neighbors(dim, J, sfc):
    f = Forest.uniform(dim, J, block_dist=sfc)
    assert f.is_uniform and f.n_blocks == (2 ** J) ** dim
    hvy, lvl, ixyz, tc = f.active(0)
    nb = f.neighbors(0)
    pos = {tuple(ixyz[k]): int(hvy[k]) for k in range(len(hvy))}
    n = 2 ** J
    n_rel = 26 if dim == 3 else 8
    for k in range(len(hvy)):
        found = 0
        for s in range(1, 57):
            lgt = nb[s - 1, hvy[k] - 1]
            if lgt < 0:
                continue
            d = _dir_of_slot(s)
            q = tuple(int((ixyz[k, a] + d[a]) % n) if a < dim else 0 for a in range(3))
            assert pos[q] == lgt, (k, s)
            found += 1
        assert found == n_rel
        assert (nb[56:, hvy[k] - 1] == -1).all()
    # treecode: digit bit0 -> y, bit1 -> x, bit2 -> z, coarsest digit highest (module_treelib.f90:793-871)
    lib = _native.host_lib()
    for k in (0, len(hvy) // 2, len(hvy) - 1):
        out = (C.c_int32 * 3)()
        lib.whost_decode(dim, J, J, int(tc[k]), out)
        assert list(out)[:dim] == list(ixyz[k])[:dim]
    one = (C.c_int32 * 3)(0, 1, 0)
    assert lib.whost_encode(3, 1, 1, one) == 1
    one = (C.c_int32 * 3)(1, 0, 0)
    assert lib.whost_encode(3, 1, 1, one) == 2
    one = (C.c_int32 * 3)(0, 0, 1)
    assert lib.whost_encode(3, 1, 1, one) == 4


def test_sfc_is_a_curve():
    """consecutive blocks along the Hilbert curve are face neighbours; every rank gets a contiguous chunk"""
    f = Forest.uniform(3, 3, block_dist="sfc_hilbert", n_ranks=4)
    tot = 0
    for r in range(4):
        hvy, lvl, ixyz, tc = f.active(r)
        tot += len(hvy)
        assert len(hvy) == 128 and (hvy == np.arange(1, 129)).all()
        step = np.abs(np.diff(ixyz, axis=0)).sum(axis=1)
        assert (step == 1).all()
    assert tot == 512


def test_hilbert_partition_matches_the_reference_fixture_files():
    """the per-rank block lists (`procs`) of files the reference wrote with 4 / 8 MPI ranks -- 2-D adaptive cylinder
    runs, 3-D equidistant and adaptive grids -- equal the contiguous chunks of our space-filling-curve order (treecode_to_hilbertcode_2D/3D,
    balanceLoad_tree.f90:203-285, 600-715).  Fixture: tests/golden/partition_procs.npz (tests/golden/make_partition_golden.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "partition_procs.npz"))
    lib = _native.host_lib()
    for k, name in enumerate(g["files"]):
        dim, Jmax = int(g[f"f{k}_dim"][0]), int(g[f"f{k}_Jmax"][0])
        procs, level, tc = g[f"f{k}_procs"], g[f"f{k}_level"], g[f"f{k}_treecode"]
        P = int(procs.max()) + 1
        ixyz = np.zeros((len(level), 3), np.int32)
        for i in range(len(level)):
            buf = (C.c_int32 * 3)()
            lib.whost_decode(dim, int(level[i]), Jmax, int(tc[i]), buf)
            ixyz[i] = list(buf)
        f = Forest.from_blocks(dim, Jmax, level, ixyz, block_dist="sfc_hilbert", n_ranks=P)
        want = {(int(l), int(t)): int(r) for l, t, r in zip(level, tc, procs)}
        for r in range(P):
            _, lv, _, t = f.active(r)
            got = {want[(int(l), int(tt))] for l, tt in zip(lv, t)}
            assert got == {r}, (str(name), r, got)


def test_two_level_forest_slots():
    """2-level grid: one level-1 block refined.  Checks the +56 / +112 slot groups are mutually consistent."""
    lv, ix = [], []
    for z in range(2):
        for y in range(2):
            for x in range(2):
                if (x, y, z) == (0, 0, 0):
                    for c in range(8):
                        lv.append(2); ix.append((c & 1, (c >> 1) & 1, (c >> 2) & 1))
                else:
                    lv.append(1); ix.append((x, y, z))
    f = Forest.from_blocks(3, 2, lv, ix)
    assert not f.is_uniform and f.n_blocks == 15
    hvy, lvl, ixyz, tc = f.active(0)
    nb = f.neighbors(0)
    by_id = {int(hvy[k]): (int(lvl[k]), tuple(ixyz[k])) for k in range(len(hvy))}
    for k in range(len(hvy)):
        for s in range(1, 169):
            lgt = nb[s - 1, hvy[k] - 1]
            if lgt < 0:
                continue
            l2, _ = by_id[int(lgt)]
            grp = (s - 1) // 56
            assert l2 - lvl[k] == (0, -1, +1)[grp]
    # the coarse block at (1,0,0) sees 4 finer neighbours across its -x face (slots 113..116) and, periodic, across +x
    kc = [k for k in range(len(hvy)) if lvl[k] == 1 and tuple(ixyz[k]) == (1, 0, 0)][0]
    assert (nb[112:116, hvy[kc] - 1] > 0).all() and (nb[116:120, hvy[kc] - 1] > 0).all()
    assert nb[0, hvy[kc] - 1] == -1


@pytest.mark.parametrize("world", [2, 5])
def test_global_refine_and_coarsen_match_the_single_rank_bookkeeping(world):
    """whost_refine_global / whost_coarsen_global on a grid partitioned over `world` ranks return, in global space-filling-curve positions,
    exactly the lists the single-rank routines return in hvy ids (one rank: hvy id == global position), and the same new grid."""
    from util import graded_blocks
    lv, ix = graded_blocks(3, 1, 3, seed=9, frac=0.3)
    n = len(lv)
    fw = Forest.from_blocks(3, 4, lv, ix, n_ranks=world, max_blocks=8 * n)
    f1 = Forest.from_blocks(3, 4, lv, ix, n_ranks=1, max_blocks=8 * n)

    def global_blocks(f):
        l, x = [], []
        for r in range(f.n_ranks):
            _, a, b, _ = f.active(r)
            l.append(a)
            x.append(b)
        return np.concatenate(l), np.concatenate(x)

    assert all(np.array_equal(a, b) for a, b in zip(global_blocks(fw), global_blocks(f1)))
    rng = np.random.default_rng(0)
    flags = (rng.random(n) < 0.4).astype(np.int32)
    new_w, *lists_w = fw.refine_global(flags)
    new_1, *lists_1 = f1.refine(flags, max_blocks=8 * n)
    assert new_w.n_ranks == world and new_w.n_blocks == new_1.n_blocks
    assert all(np.array_equal(a, b) for a, b in zip(global_blocks(new_w), global_blocks(new_1)))
    assert all(np.array_equal(a, b) for a, b in zip(lists_w, lists_1))
    counts = [new_w.n_active(r) for r in range(world)]
    assert max(counts) - min(counts) <= 1                                  # balanceLoad_tree: contiguous, equal chunks
    st = np.where(rng.random(new_w.n_blocks) < 0.8, -1, 0).astype(np.int32)
    c_w, st_w, *cl_w = new_w.coarsen_global(st, Jmin=1)
    c_1, st_1, *cl_1 = new_1.coarsen(st, Jmin=1, max_blocks=8 * n)
    assert np.array_equal(st_w, st_1) and (st_w == -1).any() and c_w.n_blocks == c_1.n_blocks < new_w.n_blocks
    assert all(np.array_equal(a, b) for a, b in zip(global_blocks(c_w), global_blocks(c_1)))
    assert all(np.array_equal(a, b) for a, b in zip(cl_w, cl_1))


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_full_tree_light_data_matches_the_oracle(seed):
    """wabbit_b200/fulltree.py (vectorised host logic of the lifted adapt_tree): tree of leaves + ancestors, treecodes and the grid decision
    (respectJmaxJmin_tree + ensureGradedness_tree with check_daughters) against the oracle's block-by-block restatement"""
    import fulltree as OFT
    import oracle as O
    from util import graded_blocks
    from wabbit_b200.fulltree import FullTree, _encode_treecodes

    class FakeSol:
        max_blocks = 100000

        class params:
            Bs = (16, 16, 16)
            wavelet = "CDF44"
            discretization = "FD_4th_central"
            useCoarseExtension = -1

        def wavelet_filter_width(self):
            return 6

    lv, ix = graded_blocks(3, 1, 4, seed, 0.25)
    forest = Forest.from_blocks(3, 4, lv, ix, max_blocks=4 * len(lv))
    ft = FullTree(FakeSol(), forest, Jmin=1)
    hvy, l, x, tc = forest.active(0)
    grid = O.Grid(level=l.astype(np.int64), ixyz=x.astype(np.int64), dim=3)
    t = OFT.Tree(O.Params(dim=3, Bs=(16,) * 3, g=6, n_eqn=1, Jmax=4), O.setup_wavelet("CDF44"), grid, np.zeros((grid.n, 1, 1, 1, 1)), Jmin=1)
    assert set(ft.slot) == set(t.blk) and ft.leaf == t.leaf and not ft.leaf_first
    i = ft._find(l.astype(np.int64), x.astype(np.int64))
    assert (i >= 0).all() and np.array_equal(ft.slots[i], hvy) and np.array_equal(_encode_treecodes(3, ft.level[i], ft.pos[i], 4), tc)
    # the neighbour rows of the tree passes (whost_ft_rows) against their numpy formulation; a second tree reuses the solver's buffer
    sol = ft.sol
    sol.set_treecodes = lambda *a: None
    ft._upload_rows()
    assert ft._rows.shape[1] == sol.max_blocks and np.array_equal(ft._rows[:, :int(ft.slots.max())], ft._rows_numpy(int(ft.slots.max())))
    assert (ft._rows[:, int(ft.slots.max()):] == -1).all() and (ft._rows[:56] >= 0).any() and (ft._rows[56:112] >= 0).any()
    keep = np.ones(len(ft.code), bool)
    keep[np.flatnonzero(ft.is_leaf)[::3]] = False                      # some other tree state on the same solver
    ft2 = FullTree(sol, forest, Jmin=1)
    ft2.code, ft2.level, ft2.pos, ft2.slots, ft2.is_leaf = (a[keep] for a in (ft.code, ft.level, ft.pos, ft.slots, ft.is_leaf))
    ft2._build_tables()
    ft2.is_leaf = ft2.child[:, 0] < 0
    ft2._upload_rows()
    assert ft2._rows is ft._rows and np.array_equal(ft2._rows[:, :int(ft2.slots.max())], ft2._rows_numpy(int(ft2.slots.max())))
    st0 = np.where(np.random.default_rng(seed).random(len(ft.code)) < 0.75, -1, 0).astype(np.int32)
    st = ft.decide(st0)
    od = OFT.decide(t, ft.status_dict(st0), 1)
    assert (st == -1).any()
    assert {k: v == -1 for k, v in ft.status_dict(st).items()} == {k: v == -1 for k, v in od.items()}


def test_refinement_flags_do_not_depend_on_the_partition():
    """refinementIndicator_tree("significant") + respectJmaxJmin_tree + ensureGradedness_tree on replicated light data: the flags in
    global space-filling-curve order are the same for 1, 2 and 3 ranks and equal to the oracle's restatement"""
    import adaptive as A
    import oracle as O
    from util import graded_blocks
    from wabbit_b200.timeloop import refinement_flags
    lv, ix = graded_blocks(3, 1, 4, 11, 0.3)
    rng = np.random.default_rng(3)
    f1 = Forest.from_blocks(3, 5, lv, ix, n_ranks=1, max_blocks=4 * len(lv))
    _, l1, x1, _ = f1.active(0)
    status = np.where(rng.random(len(l1)) < 0.3, 0, 9).astype(np.int32)
    ref = refinement_flags(f1, "significant", status, 5)
    assert 0 < ref.sum() < len(ref) and (ref[status == 9] == 1).any()          # gradedness promoted some insignificant blocks
    for world in (2, 3):
        fw = Forest.from_blocks(3, 5, lv, ix, n_ranks=world, max_blocks=4 * len(lv))
        assert np.array_equal(refinement_flags(fw, "significant", status, 5), ref)
    # the oracle's formulation (from the coarse block's point of view)
    po = O.Params(dim=3, Bs=(16, 16, 16), g=3, Jmax=5)
    grid = O.Grid(level=l1.astype(np.int64), ixyz=x1.astype(np.int64), dim=3)
    run = A.AdaptiveRun(po, "CDF40", grid, np.zeros((grid.n, 1, 1, 1, 1)), 0.0, 0, 1e-3, refinement_indicator="significant")
    run.status = status.astype(np.int64)
    assert np.array_equal(run.refine_flags("significant"), ref)
    assert np.array_equal(refinement_flags(f1, "everywhere", None, 3), (l1 < 3).astype(np.int32))


def test_host_mask_generators_match_the_oracle():
    """wabbit_b200.mask (product host code, vectorised over blocks) against the oracle's per-block restatement of create_mask_2D_ACM /
    draw_sphere: the six mask components bit for bit, the threshold_mask indicator identical"""
    import adaptive as A
    import cylinder_case as CC
    import oracle as O
    import sphere_case as SC
    from wabbit_b200 import Params
    from wabbit_b200.mask import CylinderMask2D, SphereMask3D
    rng = np.random.default_rng(0)
    # 2-D cylinder + p-norm sponge (acm_cyl.ini)
    p = Params(wavelet="CDF44", **CC.INI).finalize()
    po = O.Params(skew=False, **CC.INI)
    m, mo = CylinderMask2D(p), A.CylinderMask2D(po)
    lv = rng.integers(1, 7, 120)
    pos = np.stack([np.append(rng.integers(0, 2 ** l, 2), 0) for l in lv])
    pos[:60] = np.stack([[(2 ** l) // 2 - rng.integers(0, 2), (2 ** l) // 2 - rng.integers(0, 2), 0] for l in lv[:60]])   # around the cylinder
    f, k = m.fill(lv, pos), m.keeps(lv, pos)
    for i, (l, x) in enumerate(zip(lv, pos)):
        assert np.array_equal(f[i], mo.block(int(l), x))
        assert bool(k[i]) == mo.keeps(int(l), x)
    assert 0 < k.sum() < len(k)
    # 3-D translating sphere
    p3 = Params(wavelet="CDF44", skew_symmetry=True, **SC.INI).finalize()
    po3 = O.Params(skew=True, **SC.INI)
    s, so = SphereMask3D(p3, **SC.SPHERE), A.SphereMask3D(po3, **SC.SPHERE)
    assert s.h == so.h
    lv = rng.integers(1, 5, 200)
    pos = np.stack([rng.integers(0, 2 ** l, 3) for l in lv])
    pos[:120] = np.stack([np.clip((np.array(SC.SPHERE["center"]) * 2 ** l).astype(int) + rng.integers(-1, 2, 3), 0, 2 ** l - 1) for l in lv[:120]])
    for t in (0.0, 0.137):
        k = s.keeps(lv, pos, t)
        ko = np.array([so.keeps(int(l), x, t) for l, x in zip(lv, pos)])
        assert np.array_equal(k, ko)
        assert 0 < k.sum() < len(k)


def test_full_tree_adapt_host_path_runs_without_a_device():
    """the host side of FullTree.adapt (tree tables, neighbour rows through whost_ft_rows, pass topologies, grid decision, pruning, new
    forest, move lists) driven end to end with a device stand-in that records the calls and answers the flag queries at random: the
    light-data logic must produce a complete, graded leaf grid and consistent id lists whatever the flags are"""
    from util import graded_blocks
    from wabbit_b200.fulltree import FullTree

    class NullSol:
        max_blocks = 20000

        class params:
            Bs = (16, 16, 16)
            wavelet = "CDF44"
            discretization = "FD_4th_central"
            useCoarseExtension = -1
            n_eqn = 4
            eps = 1e-3

        def __init__(self):
            self.rng = np.random.default_rng(4)
            self.calls = []
            self.hvy_active = np.zeros(0, np.int32)

        def wavelet_filter_width(self):
            return 6

        def set_grid(self, hvy, level, tc, active=None):
            assert len(hvy) == len(level) == len(tc) and hvy.min() >= 1 and hvy.max() <= self.max_blocks
            assert len(set(hvy.tolist())) == len(hvy)
            self.known = set(hvy.tolist())
            self.calls.append("grid")
            self.set_active(hvy if active is None else active)

        def set_active(self, hvy):
            assert set(np.asarray(hvy).tolist()) <= self.known
            self.hvy_active = np.asarray(hvy)
            self.calls.append("active")

        def set_forest(self, forest, rank=0):
            self.hvy_active = forest.active(0)[0]
            self.calls.append("forest")

        def waveletDecomposition_tree(self, *a):
            self.calls.append("fwt")

        def coarse_extension_modify(self, *a):
            self.calls.append("ce")

        def waveletReconstruction_CE(self, *a):
            self.calls.append("iwt_ce")

        def coarsen_blocks(self, mothers, daughters, *a):
            assert len(daughters) == 8 * len(mothers) and len(set(daughters.tolist())) == len(daughters)
            self.calls.append("d2m")

        def threshold_tree(self, *a, **k):
            n = len(self.hvy_active)
            return np.where(self.rng.random(n) < 0.97, -1, 0).astype(np.int32), self.rng.random((n, 4))

        def patch_details(self, ids, dirs, *a, **kw):
            return self.rng.random((len(ids), 4)) * 1.1e-3          # a few pairs exceed eps * norm: the security zone acts

        def move_blocks(self, src, dst):
            assert len(src) == len(dst) and len(set(dst.tolist())) == len(dst) and len(set(src.tolist())) == len(src)
            self.moved = (np.asarray(src), np.asarray(dst))
            self.calls.append("move")

        def synchronize(self):
            pass

    lv, ix = graded_blocks(3, 1, 4, 9, 0.3)
    forest = Forest.from_blocks(3, 4, lv, ix, max_blocks=NullSol.max_blocks)
    sol = NullSol()
    ft = FullTree(sol, forest, Jmin=1)
    new, info = ft.adapt(eps=1e-3, norm=np.ones(4), use_security_zone=True, want_info=False,
                         mask_keeps=lambda level, pos: (np.asarray(pos)[:, 0] == 0))
    assert {"grid", "active", "fwt", "ce", "d2m", "move", "forest"} <= set(sol.calls)
    hvy, l, x, _ = new.active(0)
    assert 8 <= new.n_blocks < forest.n_blocks and np.array_equal(np.sort(sol.moved[1]), hvy)
    # complete and graded: every point of the domain is covered exactly once, neighbouring leaves differ by at most one level
    vol = (1.0 / 8.0 ** l.astype(np.float64)).sum()
    assert abs(vol - 1.0) < 1e-12
    nb = new.neighbors(0)
    assert ((nb[:56] >= 1).sum(axis=0) + (nb[56:112] >= 1).any(axis=0) + (nb[112:] >= 1).any(axis=0) > 0).all()
    assert len(ft.leaf_status) == new.n_blocks and set(np.unique(ft.leaf_status)) <= {0, 9}
    # the blocks the mask indicator named never disappear below their level: a kept block's position is still covered at >= its level
    kept = {(int(a), int(b[0]), int(b[1]), int(b[2])) for a, b in zip(l, x)}
    _, l0, x0, _ = forest.active(0)
    for a, b in zip(l0, x0):
        if b[0] == 0:
            assert (int(a), int(b[0]), int(b[1]), int(b[2])) in kept


def test_params_and_mask_from_the_reference_ini_format(tmp_path):
    """the .ini interface of the path: Params.from_ini + mask_from_ini on a parameter file in the reference's format (the keys of
    TESTING/acm/acm_CDF44/acm_cyl.ini that the path reads; where the reference checkout exists, that very file)"""
    import cylinder_case as CC
    from wabbit_b200 import Params
    from wabbit_b200.mask import CylinderMask2D, mask_from_ini
    text = """
[Domain]
dim=2;
domain_size=20 20;
periodic_BC=1 1;
[Blocks]
number_block_nodes=26;
number_ghost_nodes=;
number_equations=3;
eps=1.0e-3;
max_treelevel=6;
min_treelevel=;
adapt_tree=1;
refinement_indicator=everywhere;
eps_normalized=1;
eps_norm=Linfty;
threshold_mask=1;
force_maxlevel_dealiasing=1;
[Wavelet]
wavelet=CDF44;
[Time]
time_max=0.1;
CFL=1.5;
CFL_eta=0.99;
write_method=fixed_time;
write_time=0.05;
[ACM-new]
c_0=12.5;
nu=0.0;
gamma_p=0;
u_mean_set=0.0 -1.0 0.0;
inicond=meanflow;
[Sponge]
use_sponge=1;
sponge_type=p-norm;
p_sponge=8.0;
L_sponge=2.0;
C_sponge=8.0e-3;
[Discretization]
order_discretization=FD_4th_central;
[VPM]
penalization=1;
smoothing_type=cosine; hester, discontinuous/dis, cosine/cos
C_eta=1.34e-3;
geometry=cylinder;
x_cntr=10.0 10.0 0;
length=1.0;
"""
    paths = [str(tmp_path / "acm_cyl.ini")]
    open(paths[0], "w").write(text)
    if os.path.exists("/root/reference/TESTING/acm/acm_CDF44/acm_cyl.ini"):
        paths.append("/root/reference/TESTING/acm/acm_CDF44/acm_cyl.ini")
    for path in paths:
        p = Params.from_ini(path)
        for k, v in CC.INI.items():
            got = getattr(p, k)
            if k == "domain":
                assert tuple(got[:2]) == v[:2]
            else:
                assert (tuple(got) if isinstance(v, tuple) else got) == v, (k, got, v)
        got = (p.wavelet, p.eps, p.eps_normalized, p.eps_norm, p.Jmin, p.force_maxlevel_dealiasing, p.adapt_tree, p.refinement_indicator)
        assert got == ("CDF44", 1.0e-3, True, "Linfty", 1, True, True, "everywhere")
        assert p.useCoarseExtension == 1 and p.useSecurityZone == 1 and p.n_mask == 6 and not p.skew_symmetry
        assert p.threshold_mask and p.threshold_state_vector_component in ((), (1, 1, 1))
        m, ref = mask_from_ini(path, p), CylinderMask2D(p)
        assert isinstance(m, CylinderMask2D) and (m.c, m.R, m.h, m.L, m.ps) == ((10.0, 10.0), 0.5, ref.h, 2.0, 8.0)
        lv = np.array([5, 5, 3]); pos = np.array([[15, 15, 0], [16, 16, 0], [0, 0, 0]])
        assert np.array_equal(m.fill(lv, pos), ref.fill(lv, pos))
    # the example driver parses the same file and plans the run without touching a device
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "examples", "run_from_ini.py"), paths[0], "--plan"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "CylinderMask2D" in out.stdout and "Jmin..Jmax = 1..6" in out.stdout, out.stderr
    # a restart (read_from_files = 1): field files written by h5io, named in the .ini, read back and planned
    from wabbit_b200 import h5io
    lv = np.array([1, 1, 1, 1], np.int32)
    ix = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]])
    for name in ("ux", "uy", "p"):
        h5io.write_wabbit_field(str(tmp_path / f"{name}_000002000000.h5"), np.zeros((4, 27, 27)), lv, ix, np.arange(4), dim=2, Bs=(26, 26, 1),
                                domain=(20.0, 20.0, 1.0), time=2.0, iteration=77, max_level=6)
    rpath = str(tmp_path / "restart.ini")
    open(rpath, "w").write(text + "[Physics]\nread_from_files=1;\ninput_files=ux_000002000000.h5 uy_000002000000.h5 p_000002000000.h5;\n")
    pr = Params.from_ini(rpath)
    assert pr.read_from_files and pr.input_files == ("ux_000002000000.h5", "uy_000002000000.h5", "p_000002000000.h5") and pr.adapt_inicond
    out = subprocess.run([sys.executable, os.path.join(root, "examples", "run_from_ini.py"), rpath, "--plan"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "t = 2.0, iteration = 77, 4 blocks" in out.stdout, out.stdout + out.stderr


def test_refinement_flags_2d_on_the_adapted_three_vortices_grid():
    """the "significant" refinement flags on the 2-D grid the reference stores after adapt_inicond (tests/golden), statuses as stored:
    equal to the oracle's restatement of refinementIndicator_tree + ensureGradedness_tree"""
    import adaptive as A
    import adaptive_case as AC
    import oracle as O
    from wabbit_b200.timeloop import refinement_flags
    gd = AC.gold("CDF42")
    lv, ixy, status = gd["t10_level"], gd["t10_ixy"], gd["t10_status"]
    ixyz = np.concatenate([ixy, np.zeros((len(ixy), 1), np.int32)], axis=1)
    f = Forest.from_blocks(2, 4, lv.astype(np.int32), ixyz.astype(np.int32), max_blocks=400)
    _, l, x, _ = f.active(0)
    at = {(int(a), int(b[0]), int(b[1])): i for i, (a, b) in enumerate(zip(lv, ixy))}
    st = np.array([status[at[(int(a), int(b[0]), int(b[1]))]] for a, b in zip(l, x)])
    got = refinement_flags(f, "significant", st, 4)
    grid = O.Grid(level=l.astype(np.int64), ixyz=x.astype(np.int64), dim=2)
    po = O.Params(dim=2, Bs=(32, 32, 1), g=4, n_eqn=3, Jmax=4)
    run = A.AdaptiveRun(po, "CDF42", grid, np.zeros((grid.n, 1, 1, 1, 1)), 10.0, 0, 1e-3, refinement_indicator="significant")
    run.status = st.astype(np.int64)
    assert np.array_equal(got, run.refine_flags("significant")) and 0 < got.sum() < len(got)


@pytest.mark.parametrize("dim,world,seed", [(3, 2, 1), (3, 3, 5), (3, 8, 2), (2, 4, 3)])
def test_halo_plan_from_positions_equals_the_plan_from_the_neighbour_tables(dim, world, seed):
    """whost_halo_plan (block positions, the rank's own blocks only, relations taken as symmetric) against the plan read off the 168-slot
    hvy_neighbor tables of ALL ranks: halo blocks, send lists per peer, and the finer-neighbour sub-lists."""
    from util import graded_blocks
    from wabbit_b200.multi import HaloPlan, HaloPlanFromTables
    lv, ix = graded_blocks(dim, 2, 4 if dim == 3 else 5, seed, 0.3)
    forest = Forest.from_blocks(dim, 5, lv, ix, n_ranks=world, max_blocks=len(lv))
    for r in range(world):
        a, b = HaloPlan(forest, r, world), HaloPlanFromTables(forest, r, world)
        assert np.array_equal(a.halo_lgt, b.halo_lgt) and np.array_equal(a.halo_hvy, b.halo_hvy)
        assert a.recv_counts == b.recv_counts and a.send_counts == b.send_counts
        assert np.array_equal(a.send_hvy, b.send_hvy)
        assert np.array_equal(a.fine_lgt, b.fine_lgt) and np.array_equal(a.fine_recv_hvy, b.fine_recv_hvy)
        assert a.fine_recv_counts == b.fine_recv_counts and a.fine_send_counts == b.fine_send_counts
        assert np.array_equal(a.fine_send_hvy, b.fine_send_hvy)


def test_ini_keys_of_the_chebychev_integrator_and_the_filter(tmp_path):
    """[Time] time_step_method / s / RKC_custom_scheme + rows, [Discretization] filter_* (ini_file_to_params.f90:176-184, 592, 627-636)"""
    rows = {"RKC_mu": "0.0 0.5 1.6267817221296652 1.3145584466517715", "RKC_mu_tilde": "0.288421052631579 0.1442105263157895 0.4691980966984508 0.3791463309290373",
            "RKC_nu": "0.0 -1.0 -0.0770074187990374 -0.2024615056742501", "RKC_gamma_tilde": "0.0 -0.0 -0.2790201699301438 -0.1583434112210081",
            "RKC_c": "0.288421052631579 0.288421052631579 0.6371654626762986 1.0"}
    ini = tmp_path / "p.ini"
    ini.write_text("[Domain]\ndim=3;\n[Blocks]\nnumber_block_nodes=16;\nnumber_equations=4;\n[Time]\ntime_step_method=RungeKuttaChebychev;\ns=4;\n"
                   "RKC_custom_scheme=1;\n" + "".join(f"{k}={v};\n" for k, v in rows.items()) +
                   "[Discretization]\norder_discretization=FD_4th_central;\nfilter_type=explicit_5pt;\nfilter_freq=10;\nfilter_component=1 1 1 0;\n")
    p = Params.from_ini(str(ini))
    assert p.time_step_method == "RungeKuttaChebychev" and p.rkc_s == 4 and p.RKC_custom_scheme
    mu, mut, nu, gt, c = p.rkc_coefficients()
    assert len(mu) == 4 and mu[1] == 0.5 and c[-1] == 1.0 and mut[0] == c[0]
    assert p.filter_type == "explicit_5pt" and p.filter_freq == 10 and p.filter_component == (1, 1, 1, 0) and not p.filter_only_maxlevel
    q = Params()
    assert q.time_step_method == "RungeKuttaGeneric" and q.filter_type == "no_filter"
    # without a custom scheme: the tabulated scheme in closed form; the rows of the file above are its s = 4 row, as the reference prints them
    for mine, theirs in zip(q.rkc_coefficients(), (mu, mut, nu, gt, c)):
        assert np.abs(mine - theirs).max() <= 1e-15


def test_t_files_have_the_reference_row_format(tmp_path):
    """module_t_files.f90:177-201: '(n-1 (es15.8,";"), es15.8)' per row; the statistics trigger of main.f90:392"""
    from wabbit_b200 import tfiles
    assert tfiles.format_row([0.5, -1.25e-3, 1.0e10]) == " 5.00000000E-01;-1.25000000E-03; 1.00000000E+10"
    p = Params(dim=3, n_eqn=4, nu=1.0e-2, c0=10.0)
    p.penalization, p.C_eta = True, 1.0e-3
    names = ("meanflow_x", "meanflow_y", "meanflow_z", "e_kin", "ACM_energy", "mask_volume", "sponge_volume", "penal_power_solid_input",
             "penal_power_solid_dissipation", "penal_power_sponge", "force_x", "force_y", "force_z", "umag", "div_max", "div_min", "u_residual_x",
             "u_residual_y", "u_residual_z", "enstrophy", "max_vort", "helicity", "dissipation")
    stats = {k: float(i + 1) for i, k in enumerate(names)}
    for t in (0.0, 0.1):
        tfiles.write_statistics_acm(stats, t, 1.0e-3, p, 0.05, str(tmp_path))
    files = sorted(f.name for f in tmp_path.iterdir())
    assert files == sorted(["umag.t", "CFL.t", "meanflow.t", "div.t", "forces.t", "mask_volume.t", "penal_power.t", "u_residual.t", "e_kin.t",
                            "enstrophy.t", "helicity.t", "dissipation.t"])
    rows = (tmp_path / "enstrophy.t").read_text().splitlines()
    assert rows == [" 0.00000000E+00; 2.00000000E+01; 2.10000000E+01", " 1.00000000E-01; 2.00000000E+01; 2.10000000E+01"]
    umag = [float(x) for x in (tmp_path / "umag.t").read_text().splitlines()[0].split(";")]
    assert umag[1] == np.sqrt(14.0).round(8) or abs(umag[1] - np.sqrt(14.0)) < 1e-8
    assert abs(umag[4] - (np.sqrt(14.0) + np.sqrt(100.0 + 14.0))) < 1e-7
    cfl = [float(x) for x in (tmp_path / "CFL.t").read_text().splitlines()[0].split(";")]
    assert abs(cfl[2] - 1.0e-3 * 1.0e-2 / 0.05 ** 2) < 1e-10 and abs(cfl[3] - 1.0) < 1e-8
    assert tfiles.statistics_due(10, 0.33, 5, 9999999.9) and not tfiles.statistics_due(11, 0.33, 5, 9999999.9)
    assert tfiles.statistics_due(11, 0.4, 99999999, 0.2) and not tfiles.statistics_due(11, 0.41, 99999999, 0.2)


def test_is_it_time_to_save_data():
    """LIB/IO/save_data.f90:255-288 (the filter also runs right before data are saved, main.f90:370)"""
    p = Params(write_method="fixed_freq")
    p.write_freq = 4
    assert p.is_it_time_to_save_data(0.3, 8) and not p.is_it_time_to_save_data(0.3, 9)
    q = Params(write_method="fixed_time", write_time=0.5)
    assert q.is_it_time_to_save_data(1.5, 7) and q.is_it_time_to_save_data(0.5 - 1e-14, 7) and not q.is_it_time_to_save_data(0.7, 7)
    q.write_time_first = 1.0
    assert not q.is_it_time_to_save_data(0.5, 7) and q.is_it_time_to_save_data(1.0, 7)


def test_time_step_tree_dispatch_logic_without_a_device():
    """WabbitGPU.timeStep_tree (timeStep_tree.f90:26-58, main.f90:368-374) driven with a stand-in for the device calls: which integrator runs,
    when the filter runs (every filter_freq iterations, and right before data are saved), what an unknown method raises"""
    from wabbit_b200 import WabbitAbort
    from wabbit_b200.solver import WabbitGPU

    class Fake:
        def __init__(self, p):
            self.params, self.calls = p, []

        def RungeKuttaGeneric(self, t, it):
            self.calls.append("rk")
            return 0.25

        def RungeKuttaChebychev(self, t, it, *rows):
            assert len(rows) == 5 and all(len(r) == self.params.rkc_s for r in rows)
            self.calls.append("rkc")
            return 0.25

        def krylov_time_stepper(self, t, it, M, dynamic, thr):
            self.calls.append(("krylov", M, dynamic, thr))
            return 0.25, M, 0.0

        def filter_wrapper(self, ftype, comp, only, allbut):
            self.calls.append(("filter", ftype, comp))

    p = Params()
    f = Fake(p)
    assert WabbitGPU.timeStep_tree(f, 1.0, 7) == (1.25, 8, 0.25) and f.calls == ["rk"]
    p = Params(write_method="fixed_time", write_time=0.5)
    p.time_step_method, p.rkc_s, p.filter_type, p.filter_freq, p.filter_component = "RungeKuttaChebychev", 6, "explicit_7pt", 3, (1, 1, 0, 0)
    f = Fake(p)
    t, it = 0.0, 0
    for _ in range(4):                      # iterations 1..4, times 0.25 .. 1.0: filter at iteration 3 and at the save times 0.5 and 1.0
        t, it, _ = WabbitGPU.timeStep_tree(f, t, it)
    assert f.calls == ["rkc", "rkc", ("filter", "explicit_7pt", (1, 1, 0, 0)), "rkc", ("filter", "explicit_7pt", (1, 1, 0, 0)), "rkc",
                       ("filter", "explicit_7pt", (1, 1, 0, 0))]
    p = Params()
    p.time_step_method, p.M_krylov, p.krylov_subspace_dimension, p.krylov_err_threshold = "krylov", 9, "dynamic", 1e-5
    f = Fake(p)
    WabbitGPU.timeStep_tree(f, 0.0, 0)
    assert f.calls == [("krylov", 9, True, 1e-5)]
    p.time_step_method = "Leapfrog"
    with pytest.raises(WabbitAbort) as e:
        WabbitGPU.timeStep_tree(Fake(p), 0.0, 0)
    assert e.value.code == 19101816


REF_ACM = "/root/reference/TESTING/acm"


@pytest.mark.skipif(not os.path.isdir(REF_ACM), reason="the reference checkout exists in the build container only")
def test_every_acm_parameter_file_of_the_reference_parses():
    """Params.from_ini on all parameter files under TESTING/acm: wavelet, ghost nodes (setup_wavelet's value for the wavelet), stencil and
    block size as the directory names say; mask_from_ini gives the cylinder generator for the cylinder cases and refuses the insect"""
    import glob
    import re
    from wabbit_b200.mask import CylinderMask2D, mask_from_ini
    g_of = {"CDF20": 1, "CDF22": 2, "CDF40": 3, "CDF42": 4, "CDF44": 6, "CDF60": 5, "CDF62": 6}
    files = [f for f in sorted(glob.glob(os.path.join(REF_ACM, "**", "*.ini"), recursive=True)) if "kinematics" not in f]
    assert len(files) >= 17
    for f in files:
        p = Params.from_ini(f)
        case = os.path.basename(os.path.dirname(f))
        m = re.search(r"(CDF\d\d)", case)
        assert m and p.wavelet == m.group(1) and p.g == g_of[p.wavelet], (case, p.wavelet, p.g)
        fd = re.search(r"FD(\d)", case)
        if fd:
            assert p.discretization == {"2": "FD_2nd_central", "4": "FD_4th_central", "6": "FD_6th_central"}[fd.group(1)]
        assert p.time_step_method == "RungeKuttaGeneric" and p.filter_type in ("no_filter", "")
        if case.startswith("acm_"):
            assert p.dim == 2 and tuple(p.Bs[:2]) == (26, 26) and p.penalization and p.use_sponge
            assert isinstance(mask_from_ini(f, p), CylinderMask2D)
        elif case.startswith("bumblebee"):
            with pytest.raises(ValueError):
                mask_from_ini(f, p)
        else:
            assert mask_from_ini(f, p) is None
