"""wabbit_b200/h5io.py: WABBIT field files written without an HDF5 library.

  * message encodings byte-identical to those in the reference's own files (hex strings below were read from
    TESTING/acm/acm_CDF44/ux_000000050000.h5 with oracle/h5lite.py);
  * a file written from the state of an adaptive run reads back (oracle/h5lite.py, an independent reader developed against the
    reference's files) with every dataset and attribute intact;
  * where the reference checkout exists (the build container): a stored fixture is re-written and read back -- all eight datasets and all
    ten attributes identical to the original's.
"""
import os
import struct

import numpy as np
import pytest

import h5lite
from wabbit_b200 import h5io

REF = "/root/reference/TESTING/acm/acm_CDF44/ux_000000050000.h5"


def test_message_encodings_match_the_reference_files():
    # dataspace of `level` [124]; of `blocks` [124, 27, 27]
    assert h5io._dataspace_msg((124,)).hex() == "01010100000000007c000000000000007c00000000000000"
    assert h5io._dataspace_msg((124, 27, 27)).hex() == ("01030100000000007c000000000000001b000000000000001b00000000000000"
                                                          "7c000000000000001b000000000000001b00000000000000")
    # datatype messages (float64 of `blocks`, int32 of `level`, int64 of `block_treecode_num`), padded as stored
    assert h5io._pad8(h5io.DT_F64).hex() == "11203f000800000000004000340b0034ff03000000000000"
    assert h5io._pad8(h5io.DT_I32).hex() == "10080000040000000000200000000000"
    assert h5io._pad8(h5io.DT_I64).hex() == "10080000080000000000400000000000"
    # whole attribute messages incl. their 8-byte message headers: time = 0.05, dim = 2, block-size = (26, 26, 1)
    assert h5io._attribute_msg("time", np.array([0.05])).hex() == (
        "0c00480000000000" "010005001400180074696d650000000011203f000800000000004000340b0034ff03000000000000"
        "0101010000000000010000000000000001000000000000009a9999999999a93f")
    assert h5io._attribute_msg("dim", np.array([2], np.int32)).hex() == (
        "0c00400000000000" "010004000c00180064696d000000000010080000040000000000200000000000"
        "0101010000000000010000000000000001000000000000000200000000000000")
    assert h5io._attribute_msg("block-size", np.array([26, 26, 1], np.int32)).hex().startswith(
        "0c00500000000000" "01000b000c001800626c6f636b2d73697a6500000000000010080000040000000000200000000000"
        "0101010000000000030000000000000003000000000000001a0000001a000000")


def test_superblock_and_group_structures(tmp_path):
    path = str(tmp_path / "t.h5")
    h5io.write_h5(path, {"b": np.arange(6, dtype=np.float64).reshape(2, 3), "a": np.arange(4, dtype=np.int32)}, {"b": {"k": np.array([7], np.int32)}})
    b = open(path, "rb").read()
    assert b[:8] == b"\x89HDF\r\n\x1a\n" and list(b[8:16]) == [0, 0, 0, 0, 0, 8, 8, 0]
    assert struct.unpack_from("<HHI", b, 16) == (4, 16, 0)
    base, fs, eof, drv = struct.unpack_from("<QQQQ", b, 24)
    assert (base, fs, drv) == (0, h5io.UNDEF, h5io.UNDEF) and eof == len(b)
    assert struct.unpack_from("<QQII", b, 56) == (0, 96, 1, 0) and struct.unpack_from("<QQ", b, 80) == (136, 680)
    assert struct.unpack_from("<BBHII", b, 96) == (1, 0, 1, 1, 24)                       # root object header, as in the reference's files
    assert b[136:140] == b"TREE" and b[680:684] == b"HEAP" and b[680 + 32 + 8:680 + 32 + 10] == b"a\x00"
    seg_size, free_at, seg_addr = struct.unpack_from("<QQQ", b, 688)
    assert seg_addr == 712 and struct.unpack_from("<QQ", b, seg_addr + free_at) == (1, seg_size - free_at)    # one free block, H5HL_FREE_NULL
    f = h5lite.H5Lite(path)
    assert sorted(f.datasets) == ["a", "b"]
    assert np.array_equal(f.read("b"), np.arange(6.0).reshape(2, 3)) and np.array_equal(f.read("a"), np.arange(4))
    assert np.array_equal(f.attrs("b")["k"], [7])


def test_field_file_of_an_adaptive_grid_reads_back(tmp_path):
    gd = np.load(os.path.join(os.path.dirname(__file__), "golden", "three_vortices_adapt_FD4_CDF40.npz"))
    level, ixy, u = gd["t10_level"], gd["t10_ixy"], gd["t10_u"]
    nb, Bs = len(level), 32
    ixyz = np.concatenate([ixy, np.zeros((nb, 1), np.int32)], axis=1)
    field = np.zeros((nb, Bs + 1, Bs + 1))
    field[:, :Bs, :Bs] = u[:, 0]
    tc = np.arange(nb, dtype=np.int64) * 4
    path = str(tmp_path / "ux_000010000000.h5")
    h5io.write_wabbit_field(path, field, level, ixyz, tc, dim=2, Bs=(Bs, Bs, 1), domain=(6.283185307179586,) * 3, time=10.0, iteration=3054,
                            max_level=4, refinement_status=gd["t10_status"])
    d = h5lite.read_wabbit(path)
    assert np.array_equal(d["blocks"], field) and np.array_equal(d["level"].ravel(), level) and np.array_equal(d["treecode"].ravel(), tc)
    assert np.array_equal(d["refinement_status"].ravel(), gd["t10_status"])
    got_ixy = np.rint(d["origin"][:, ::-1] / (d["spacing"][:, ::-1] * Bs)).astype(np.int32)
    assert np.array_equal(got_ixy, ixy)
    a = d["attrs"]
    assert a["version"][0] == 20240410 and a["dim"][0] == 2 and list(a["block-size"]) == [32, 32, 1] and a["time"][0] == 10.0
    assert a["iteration"][0] == 3054 and a["total_number_blocks"][0] == nb and a["max_level"][0] == 4
    assert np.array_equal(a["domain-size"], [6.283185307179586] * 2) and list(a["periodic_BC"]) == [1, 1, 1]


@pytest.mark.skipif(not os.path.exists(REF), reason="reference checkout not present")
def test_rewriting_a_reference_fixture_preserves_everything(tmp_path):
    src = h5lite.H5Lite(REF)
    o = h5lite.read_wabbit(REF)
    a = o["attrs"]
    Bs = [int(v) for v in a["block-size"]]
    dim = int(a["dim"][0])
    ixy = np.rint(o["origin"][:, ::-1] / (o["spacing"][:, ::-1] * Bs[0])).astype(np.int64)
    ixyz = np.concatenate([ixy, np.zeros((len(ixy), 1), np.int64)], axis=1)
    path = str(tmp_path / "copy.h5")
    h5io.write_wabbit_field(path, o["blocks"], o["level"].ravel(), ixyz, o["treecode"].ravel(), dim=dim, Bs=Bs,
                            domain=tuple(a["domain-size"]) + (1.0,), time=float(a["time"][0]), iteration=int(a["iteration"][0]),
                            max_level=int(a["max_level"][0]), refinement_status=o["refinement_status"].ravel(),
                            periodic=a["periodic_BC"], symmetry=a["symmetry_BC"], lgt_ids=src.read("lgt_ids").ravel(), procs=src.read("procs").ravel())
    new = h5lite.H5Lite(path)
    assert sorted(new.datasets) == sorted(src.datasets)
    for name in src.datasets:
        x, y = src.read(name), new.read(name)
        assert x.dtype == y.dtype and np.array_equal(x.reshape(y.shape), y), name      # origin / spacing recomputed from the treecode geometry: bit-equal
    # the numerical treecodes of the host forest are the file's block_treecode_num
    from wabbit_b200 import Forest
    f = Forest.from_blocks(dim, int(a["max_level"][0]), o["level"].ravel().astype(np.int32), ixyz.astype(np.int32), max_blocks=len(ixyz) + 8)
    _, fl, fx, ftc = f.active(0)
    ref_tc = {(int(L), int(q[0]), int(q[1])): int(t) for L, q, t in zip(o["level"].ravel(), ixy, o["treecode"].ravel())}
    assert all(ref_tc[(int(L), int(q[0]), int(q[1]))] == int(t) for L, q, t in zip(fl, fx, ftc))
    an = new.attrs("blocks")
    assert sorted(an) == sorted(a)
    for k in a:
        assert a[k].dtype == an[k].dtype and np.array_equal(a[k], an[k]), k


def test_save_data_of_an_oracle_run_equals_the_reference_output(tmp_path):
    """the state after adapt_inicond of the 3vortices case (oracle, CPU) saved with save_data: file names as the reference's, grid, statuses
    and fields equal to the fixture the reference wrote (which stores the first upper ghost point of every block as well)"""
    import adaptive_case as AC
    from test_oracle_adaptive import make_run
    from wabbit_b200 import Forest, Params
    run = make_run("CDF40")
    p = Params(wavelet="CDF40", g=3, skew_symmetry=True, **AC.INI).finalize()
    forest = Forest.from_blocks(2, p.Jmax, run.grid.level.astype(np.int32), run.grid.ixyz.astype(np.int32), max_blocks=400)
    hvy, lvl, pos, tc = forest.active(0)
    at = {(int(l), int(x[0]), int(x[1])): b for b, (l, x) in enumerate(zip(run.grid.level, run.grid.ixyz))}
    o = np.array([at[(int(l), int(x[0]), int(x[1]))] for l, x in zip(lvl, pos)])
    paths = h5io.save_data(str(tmp_path), ("ux", "uy", "p"), run.u[o], lvl, pos, tc, p, run.time, run.iteration, refinement_status=run.status[o])
    assert [os.path.basename(q) for q in paths] == ["ux_000010000000.h5", "uy_000010000000.h5", "p_000010000000.h5"]
    gd = AC.gold("CDF40")
    fields = []
    for q in paths:
        d = h5lite.read_wabbit(q)
        fields.append(d["blocks"])
    ixy = np.rint(d["origin"][:, ::-1] / (d["spacing"][:, ::-1] * AC.BS)).astype(np.int64)
    u = np.stack(fields, axis=1)[:, :, :AC.BS, :AC.BS]
    err = AC.compare(gd, "t10", d["level"].ravel(), ixy, d["refinement_status"].ravel(), u, int(d["attrs"]["iteration"][0]), float(d["attrs"]["time"][0]))
    assert err <= 1e-15
    assert np.array_equal(d["treecode"].ravel(), tc)


def test_product_reader_round_trip_and_reference_files(tmp_path):
    """wabbit_b200.h5io.read_wabbit_field / read_state (the restart side): own files round-trip; where the reference checkout exists, its
    chunked files read the same as with the independent test reader"""
    rng = np.random.default_rng(0)
    nb, Bs = 7, (8, 6, 4)
    level = rng.integers(1, 4, nb).astype(np.int32)
    ixyz = np.stack([rng.integers(0, 2 ** l, 3) for l in level])
    field = rng.standard_normal((nb, Bs[2] + 1, Bs[1] + 1, Bs[0] + 1))
    tc = rng.integers(0, 1 << 40, nb)
    paths = []
    for name in ("ux", "uy"):
        path = str(tmp_path / f"{name}_000000500000.h5")
        h5io.write_wabbit_field(path, field if name == "ux" else 2 * field, level, ixyz, tc, dim=3, Bs=Bs, domain=(1.0, 2.0, 3.0), time=0.5, iteration=12,
                                max_level=5, refinement_status=np.arange(nb))
        paths.append(path)
    d = h5io.read_wabbit_field(paths[0])
    assert np.array_equal(d["blocks"], field) and np.array_equal(d["level"], level) and np.array_equal(d["ixyz"], ixyz)
    assert np.array_equal(d["treecode"], tc) and np.array_equal(d["refinement_status"], np.arange(nb)) and d["Bs"] == [8, 6, 4]
    st = h5io.read_state(paths, g=3)
    assert st["hvy"].shape == (nb, 2, 4 + 6, 6 + 6, 8 + 6) and st["time"] == 0.5 and st["iteration"] == 12
    assert np.array_equal(st["hvy"][:, 1, 3:7, 3:9, 3:11], 2 * field[:, :4, :6, :8]) and st["hvy"][:, :, :3].max() == 0.0
    if os.path.exists(REF):
        o, r = h5lite.read_wabbit(REF), h5io.read_wabbit_field(REF)
        assert np.array_equal(o["blocks"], r["blocks"]) and np.array_equal(o["level"].ravel(), r["level"])
        assert np.array_equal(o["treecode"].ravel(), r["treecode"]) and np.array_equal(o["refinement_status"].ravel(), r["refinement_status"])
        assert all(np.array_equal(o["attrs"][k], r["attrs"][k]) for k in o["attrs"])
        ref3d = "/root/reference/TESTING/acm/taylorGreen/taylorGreenEqui_FD4_CDF40/ux_000010000000.h5"
        o, r = h5lite.read_wabbit(ref3d), h5io.read_wabbit_field(ref3d)
        assert np.array_equal(o["blocks"], r["blocks"]) and r["dim"] == 3
        want = np.rint(o["origin"][:, ::-1] / (o["spacing"][:, ::-1] * r["Bs"][0])).astype(np.int64)
        assert np.array_equal(r["ixyz"], want)


@pytest.mark.skipif(not os.path.exists("/root/reference/TESTING/acm/taylorGreen"), reason="reference checkout not present")
def test_rewriting_a_3d_reference_fixture(tmp_path):
    """the 3-D layout (blocks [Nb, Bz+1, By+1, Bx+1], origin / spacing stored z, y, x): a Taylor-Green fixture re-written and read back"""
    ref = "/root/reference/TESTING/acm/taylorGreen/taylorGreenEqui_FD4_CDF40/p_000010000000.h5"
    r = h5io.read_wabbit_field(ref)
    a = r["attrs"]
    path = str(tmp_path / "p.h5")
    h5io.write_wabbit_field(path, r["blocks"], r["level"], r["ixyz"], r["treecode"], dim=3, Bs=r["Bs"], domain=tuple(a["domain-size"]),
                            time=float(a["time"][0]), iteration=int(a["iteration"][0]), max_level=int(a["max_level"][0]),
                            refinement_status=r["refinement_status"], periodic=a["periodic_BC"], symmetry=a["symmetry_BC"])
    src, new = h5lite.H5Lite(ref), h5lite.H5Lite(path)
    for name in ("blocks", "block_treecode_num", "level", "coords_origin", "coords_spacing", "refinement_status"):
        x, y = src.read(name), new.read(name)
        assert np.array_equal(x.reshape(y.shape), y), name
    an = new.attrs("blocks")
    for k in a:
        assert np.array_equal(a[k], an[k]), k
    st = h5io.read_state([path], g=3)
    assert st["hvy"].shape[1:] == (1, 26, 26, 26) and st["iteration"] == int(a["iteration"][0])
