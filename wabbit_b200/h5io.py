"""Writer and reader for WABBIT's field files: what saveHDF5_tree leaves on disk (LIB/MESH/InputOutput.f90:237-280, 322-764;
LIB/MODULE/module_hdf5_wrapper.f90), so that the reference's own tools (wabbit-post, the python-tools, ParaView readers) and a restart
(`read_from_files = 1`) can consume what the device path produced.

  datasets    blocks [Nb, (Bz+1,) By+1, Bx+1] float64 (interior + the first upper ghost point, as the reference saves them),
              block_treecode_num [Nb] int64, level [Nb] int32, coords_origin / coords_spacing [Nb, dim] float64 (stored z, y, x),
              refinement_status, lgt_ids, procs [Nb] int32
  attributes  on `blocks`: domain-size, periodic_BC, symmetry_BC, version (20240410), block-size, time, iteration,
              total_number_blocks, max_level, dim

No HDF5 library exists in this image, so the file is assembled directly in the classic on-disk format the reference's files use
(superblock version 0, version-1 object headers, a symbol-table root group with one B-tree node / local heap / symbol-table node,
version-1 dataspace and attribute messages, version-3 layout messages) -- with CONTIGUOUS instead of chunked raw data, which every HDF5
reader handles alike.  The message encodings are byte-identical to those found in the reference's own fixture files
(tests/test_h5io.py compares them); the files have been read back with the repository's independent reader (oracle/h5lite.py), not yet
with libhdf5 itself.
"""
from __future__ import annotations

import struct
from typing import Dict, Sequence

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
VERSION = 20240410

# datatype messages as libhdf5 writes them for native little-endian types (taken from the reference's files)
DT_F64 = bytes.fromhex("11203f000800000000004000340b0034ff030000")
DT_I32 = bytes.fromhex("100800000400000000002000")
DT_I64 = bytes.fromhex("100800000800000000004000")
FILL_MSG = bytes.fromhex("0201020100000000")      # version 2, early allocation, write fill value if set, defined, size 0


def _pad8(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 8)


def _dtype_msg(a: np.ndarray) -> bytes:
    if a.dtype == np.float64:
        return DT_F64
    if a.dtype == np.int32:
        return DT_I32
    if a.dtype == np.int64:
        return DT_I64
    raise TypeError(f"unsupported dtype {a.dtype}")


def _dataspace_msg(shape: Sequence[int]) -> bytes:
    """version 1, maximum dimensions present (= the dimensions)"""
    dims = b"".join(struct.pack("<Q", int(n)) for n in shape)
    return struct.pack("<BBB5x", 1, len(shape), 1) + dims + dims


def _message(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _attribute_msg(name: str, value: np.ndarray) -> bytes:
    nm = name.encode() + b"\x00"
    dt = _dtype_msg(value)
    ds = _dataspace_msg(value.shape)
    body = struct.pack("<BxHHH", 1, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + value.tobytes()
    return _message(0x000C, body)


def _object_header(messages: Sequence[bytes]) -> bytes:
    body = b"".join(messages)
    return struct.pack("<BxHII4x", 1, len(messages), 1, len(body)) + body


def write_h5(path: str, datasets: Dict[str, np.ndarray], attrs: Dict[str, Dict[str, np.ndarray]]) -> None:
    """Write little-endian float64 / int32 / int64 arrays as contiguous datasets of the root group; attrs[name] are the attributes of
    dataset `name`.  At most 8 datasets (one symbol-table node of the default group size, as in WABBIT's files)."""
    names = sorted(datasets)                                  # symbol-table entries are ordered by name
    if not 0 < len(names) <= 8:
        raise ValueError("1 .. 8 datasets")
    # ---- local heap: the empty name of the root at offset 0, then the link names, then one free block
    heap_off, seg = {}, bytearray(8)
    for n in names:
        heap_off[n] = len(seg)
        seg += _pad8(n.encode() + b"\x00")
    free_at = len(seg)
    seg += struct.pack("<QQ", 1, 32) + bytes(16)              # free block: next = H5HL_FREE_NULL (1), size 32
    # ---- fixed layout of the metadata
    a_root, a_btree, a_heap = 96, 136, 680
    a_seg = a_heap + 32
    a_snod = a_seg + len(seg)
    pos = a_snod + 8 + 8 * 40
    hdrs, where = {}, {}
    for n in names:
        arr = np.ascontiguousarray(datasets[n])
        if arr.dtype.byteorder == ">":
            raise TypeError("little-endian arrays only")
        msgs = [_message(0x0001, _dataspace_msg(arr.shape)), _message(0x0003, _dtype_msg(arr), flags=1), _message(0x0005, FILL_MSG, flags=1)]
        att = [_attribute_msg(k, np.ascontiguousarray(v)) for k, v in attrs.get(n, {}).items()]
        layout_len = 8 + 24
        hdr_len = 16 + sum(len(m) for m in msgs) + layout_len + sum(len(m) for m in att)
        a_hdr = pos
        a_data = a_hdr + hdr_len
        a_data += -a_data % 8
        layout = _message(0x0008, struct.pack("<BBQQ", 3, 1, a_data, arr.nbytes))          # version 3, class 1 = contiguous
        hdrs[n] = _object_header(msgs + [layout] + att)
        assert len(hdrs[n]) == hdr_len
        where[n] = (a_hdr, a_data, arr)
        pos = a_data + arr.nbytes
        pos += -pos % 8
    eof = pos
    out = bytearray(eof)
    # ---- superblock version 0
    out[0:8] = b"\x89HDF\r\n\x1a\n"
    out[8:16] = bytes([0, 0, 0, 0, 0, 8, 8, 0])
    struct.pack_into("<HHI", out, 16, 4, 16, 0)               # group leaf node K, group internal node K, consistency flags
    struct.pack_into("<QQQQ", out, 24, 0, UNDEF, eof, UNDEF)  # base address, free-space info, end of file, driver info
    struct.pack_into("<QQII", out, 56, 0, a_root, 1, 0)       # root symbol-table entry: cached symbol-table information
    struct.pack_into("<QQ", out, 80, a_btree, a_heap)
    # ---- root group: object header with the symbol-table message
    root = _object_header([_message(0x0011, struct.pack("<QQ", a_btree, a_heap))])
    out[a_root:a_root + len(root)] = root
    # ---- B-tree node (type 0, leaf, one child): key0 = "", key1 = the largest name
    out[a_btree:a_btree + 4] = b"TREE"
    struct.pack_into("<BBHQQ", out, a_btree + 4, 0, 0, 1, UNDEF, UNDEF)
    struct.pack_into("<QQQ", out, a_btree + 24, 0, a_snod, heap_off[names[-1]])
    # ---- local heap
    out[a_heap:a_heap + 4] = b"HEAP"
    struct.pack_into("<B3xQQQ", out, a_heap + 4, 0, len(seg), free_at, a_seg)
    out[a_seg:a_seg + len(seg)] = seg
    # ---- symbol-table node
    out[a_snod:a_snod + 4] = b"SNOD"
    struct.pack_into("<BBH", out, a_snod + 4, 1, 0, len(names))
    for i, n in enumerate(names):
        struct.pack_into("<QQII", out, a_snod + 8 + 40 * i, heap_off[n], where[n][0], 0, 0)
    # ---- datasets
    for n in names:
        a_hdr, a_data, arr = where[n]
        out[a_hdr:a_hdr + len(hdrs[n])] = hdrs[n]
        out[a_data:a_data + arr.nbytes] = arr.tobytes()
    with open(path, "wb") as f:
        f.write(out)


def write_wabbit_field(path: str, field: np.ndarray, level: np.ndarray, ixyz: np.ndarray, treecode: np.ndarray, *, dim: int, Bs: Sequence[int],
                       domain: Sequence[float], time: float, iteration: int, max_level: int, refinement_status=None, periodic=(1, 1, 1),
                       symmetry=(0, 0, 0), lgt_ids=None, procs=None) -> None:
    """One field of saveHDF5_tree: `field` [Nb, (Bz+1,) By+1, Bx+1] (interior plus the first upper ghost point of every block, what a
    ghosted download holds at [g : g+Bs+1]); level / ixyz [Nb, 3] (zero-based block coordinates) / treecode [Nb] as Forest.active returns
    them.  Origin and spacing follow get_block_spacing_origin: dx = 2^-J L / Bs, x0 = ixyz Bs dx."""
    nb = len(level)
    level = np.asarray(level, dtype=np.int32)
    dx = np.stack([2.0 ** (-level.astype(np.float64)) * float(domain[d]) / float(Bs[d]) for d in range(dim)], axis=1)
    x0 = np.asarray(ixyz)[:, :dim].astype(np.float64) * np.asarray(Bs[:dim], dtype=np.float64)[None, :] * dx
    want = (nb,) + tuple(int(Bs[d]) + 1 for d in reversed(range(dim)))
    if tuple(field.shape) != want:
        raise ValueError(f"field has shape {field.shape}, expected {want}")
    bs3 = [int(Bs[d]) if d < dim else 1 for d in range(3)]
    i32 = lambda v: np.asarray(v, dtype=np.int32)
    datasets = {
        "blocks": np.ascontiguousarray(field, dtype=np.float64),
        "block_treecode_num": np.asarray(treecode, dtype=np.int64),
        "level": level,
        "coords_origin": np.ascontiguousarray(x0[:, ::-1]),          # stored (z,) y, x
        "coords_spacing": np.ascontiguousarray(dx[:, ::-1]),
        "refinement_status": i32(np.zeros(nb) if refinement_status is None else refinement_status),
        "lgt_ids": i32(np.arange(1, nb + 1) if lgt_ids is None else lgt_ids),
        "procs": i32(np.ones(nb) if procs is None else procs),
    }
    attrs = {"blocks": {
        "domain-size": np.asarray(domain[:dim], dtype=np.float64),
        "periodic_BC": i32(periodic), "symmetry_BC": i32(symmetry), "version": i32([VERSION]), "block-size": i32(bs3),
        "time": np.asarray([time], dtype=np.float64), "iteration": i32([iteration]), "total_number_blocks": i32([nb]),
        "max_level": i32([max_level]), "dim": i32([dim]),
    }}
    write_h5(path, datasets, attrs)


def save_data(directory: str, field_names: Sequence[str], hvy: np.ndarray, level, ixyz, treecode, params, time: float, iteration: int,
              refinement_status=None) -> list:
    """save_data (LIB/MESH/InputOutput.f90 / main.f90:432-440) for the state vector: one file `<name>_<nint(time*1e6):012d>.h5` per
    component.  hvy: ghosted host array [Nb, ncomp, nz, ny, nx] in the order of level / ixyz / treecode (what wgpu_download with
    g_sync >= 1 fills, ordered as Forest.active) -- the first upper ghost point of every block is saved with the interior, as the
    reference does."""
    import os
    p, g, dim = params, params.g, params.dim
    Bs = [int(v) for v in p.Bs]
    sl = (slice(g, g + Bs[1] + 1), slice(g, g + Bs[0] + 1))
    paths = []
    for c, name in enumerate(field_names):
        if dim == 3:
            field = hvy[:, c, g:g + Bs[2] + 1, sl[0], sl[1]]
        else:
            field = hvy[:, c, 0, sl[0], sl[1]]
        path = os.path.join(directory, f"{name}_{int(round(time * 1.0e6)):012d}.h5")
        write_wabbit_field(path, field, level, ixyz, treecode, dim=dim, Bs=Bs, domain=p.domain, time=time, iteration=iteration, max_level=p.Jmax,
                           refinement_status=refinement_status, periodic=p.periodic)
        paths.append(path)
    return paths


# ----------------------------------------------------------------------------------------------------------------------
# Reader (readHDF5vct_tree's file side, LIB/MESH/InputOutput.f90:322-764): restart from field files -- the reference's (chunked) or ours
# ----------------------------------------------------------------------------------------------------------------------
class _File:
    """the subset of the classic HDF5 format WABBIT's files use: superblock v0, symbol-table groups, v1 object headers (with continuation
    blocks), contiguous / compact / chunked (v1 B-tree, unfiltered) datasets, v1 attributes of integer and floating-point type"""

    def __init__(self, path: str):
        with open(path, "rb") as f:
            self.b = b = f.read()
        if b[:8] != b"\x89HDF\r\n\x1a\n" or b[8] != 0:
            raise ValueError(f"{path}: not a classic (superblock version 0) HDF5 file")
        btree, heap = struct.unpack_from("<QQ", b, 80)
        if b[heap:heap + 4] != b"HEAP":
            raise ValueError(f"{path}: root group has no local heap")
        self.names: Dict[str, int] = {}
        self._group(btree, struct.unpack_from("<Q", b, heap + 24)[0])

    def _group(self, node: int, seg: int):
        b = self.b
        if b[node:node + 4] == b"TREE":
            n = struct.unpack_from("<H", b, node + 6)[0]
            for i in range(n):
                self._group(struct.unpack_from("<Q", b, node + 32 + 16 * i)[0], seg)
        elif b[node:node + 4] == b"SNOD":
            for i in range(struct.unpack_from("<H", b, node + 6)[0]):
                off, hdr = struct.unpack_from("<QQ", b, node + 8 + 40 * i)
                self.names[b[seg + off:b.index(b"\x00", seg + off)].decode()] = hdr
        else:
            raise ValueError("corrupt group structure")

    def _msgs(self, hdr: int):
        b = self.b
        version, nmsg = b[hdr], struct.unpack_from("<H", b, hdr + 2)[0]
        if version != 1:
            raise ValueError("only version-1 object headers are supported")
        todo, out = [(hdr + 16, struct.unpack_from("<I", b, hdr + 8)[0])], []
        while todo and len(out) < nmsg:
            p, n = todo.pop(0)
            end = p + n
            while p + 8 <= end and len(out) < nmsg:
                t, size = struct.unpack_from("<HH", b, p)
                if t == 0x10:
                    todo.append(struct.unpack_from("<QQ", b, p + 8))
                out.append((t, p + 8))
                p += 8 + size
        return out

    def _dtype(self, p: int) -> np.dtype:
        cls, size = self.b[p] & 15, struct.unpack_from("<I", self.b, p + 4)[0]
        order = ">" if self.b[p + 1] & 1 else "<"
        if cls == 0:
            return np.dtype(f"{order}{'i' if self.b[p + 1] & 8 else 'u'}{size}")
        if cls == 1:
            return np.dtype(f"{order}f{size}")
        raise ValueError(f"datatype class {cls} is not supported")

    def _shape(self, p: int):
        version, rank = self.b[p], self.b[p + 1]
        return tuple(struct.unpack_from("<Q", self.b, p + (8 if version == 1 else 4) + 8 * i)[0] for i in range(rank))

    def dataset(self, name: str) -> np.ndarray:
        b = self.b
        shape = dt = lay = None
        for t, p in self._msgs(self.names[name]):
            if t == 1:
                shape = self._shape(p)
            elif t == 3:
                dt = self._dtype(p)
            elif t == 8:
                lay = p
        if b[lay] != 3:
            raise ValueError("only version-3 layout messages are supported")
        n = int(np.prod(shape))
        if b[lay + 1] == 1:
            return np.frombuffer(b, dt, n, struct.unpack_from("<Q", b, lay + 2)[0]).reshape(shape).astype(dt.newbyteorder("="))
        if b[lay + 1] == 0:
            return np.frombuffer(b, dt, n, lay + 4).reshape(shape).astype(dt.newbyteorder("="))
        rank1 = b[lay + 2]
        cdims = struct.unpack_from(f"<{rank1}I", b, lay + 11)[:-1]
        out = np.zeros(shape, dt.newbyteorder("="))
        self._chunks(struct.unpack_from("<Q", b, lay + 3)[0], rank1, cdims, dt, out)
        return out

    def _chunks(self, node: int, rank1: int, cdims, dt, out):
        b = self.b
        level, n = b[node + 5], struct.unpack_from("<H", b, node + 6)[0]
        ksz = 8 + 8 * rank1
        for i in range(n):
            k = node + 24 + i * (ksz + 8)
            nbytes, filt = struct.unpack_from("<II", b, k)
            child = struct.unpack_from("<Q", b, k + ksz)[0]
            if level:
                self._chunks(child, rank1, cdims, dt, out)
                continue
            if filt:
                raise ValueError("filtered chunks are not supported")
            offs = struct.unpack_from(f"<{rank1}Q", b, k + 8)[:-1]
            c = np.frombuffer(b, dt, nbytes // dt.itemsize, child).reshape(cdims)
            sl = tuple(slice(o, min(o + e, s)) for o, e, s in zip(offs, cdims, out.shape))
            out[sl] = c[tuple(slice(0, s.stop - s.start) for s in sl)]

    def attributes(self, name: str) -> Dict[str, np.ndarray]:
        b, res = self.b, {}
        for t, p in self._msgs(self.names[name]):
            if t != 0x0C:
                continue
            if b[p] != 1:
                raise ValueError("only version-1 attribute messages are supported")
            nsz, dsz, ssz = struct.unpack_from("<HHH", b, p + 2)
            up = lambda v: (v + 7) & ~7
            q = p + 8
            key = b[q:q + nsz].split(b"\x00")[0].decode()
            dt = self._dtype(q + up(nsz))
            shape = self._shape(q + up(nsz) + up(dsz)) if ssz >= 8 else ()
            res[key] = np.frombuffer(b, dt, int(np.prod(shape)) if shape else 1, q + up(nsz) + up(dsz) + up(ssz)).astype(dt.newbyteorder("="))
        return res


def read_wabbit_field(path: str) -> dict:
    """One field file: blocks [Nb, (Bz+1,) By+1, Bx+1], level [Nb], ixyz [Nb, 3] (zero-based block coordinates recovered from origin /
    spacing), treecode [Nb], refinement_status [Nb], and the attributes (time, iteration, block-size, domain-size, max_level, dim, ...)"""
    f = _File(path)
    a = f.attributes("blocks")
    dim = int(a["dim"][0]) if "dim" in a else f.dataset("coords_origin").shape[1]
    Bs = [int(v) for v in a["block-size"]]
    origin, spacing = f.dataset("coords_origin"), f.dataset("coords_spacing")
    ixyz = np.zeros((origin.shape[0], 3), dtype=np.int64)
    for d in range(dim):                                       # the file stores (z,) y, x
        ixyz[:, d] = np.rint(origin[:, dim - 1 - d] / (spacing[:, dim - 1 - d] * Bs[d]))
    out = {"blocks": f.dataset("blocks"), "level": f.dataset("level").ravel().astype(np.int32), "ixyz": ixyz,
           "treecode": f.dataset("block_treecode_num").ravel().astype(np.int64), "attrs": a, "dim": dim, "Bs": Bs}
    out["refinement_status"] = f.dataset("refinement_status").ravel().astype(np.int32) if "refinement_status" in f.names else None
    return out


def read_state(paths: Sequence[str], g: int) -> dict:
    """readHDF5vct_tree for a list of field files on the same grid (`input_files` of the .ini): the state vector as a ghosted host array
    [Nb, ncomp, nz, ny, nx] (interiors filled, ghost nodes zero -- to be synchronised by the consumer) + the light data of the grid"""
    first = read_wabbit_field(paths[0])
    dim, Bs, nb = first["dim"], first["Bs"], len(first["level"])
    nz = Bs[2] + 2 * g if dim == 3 else 1
    hvy = np.zeros((nb, len(paths), nz, Bs[1] + 2 * g, Bs[0] + 2 * g))
    for c, path in enumerate(paths):
        d = first if c == 0 else read_wabbit_field(path)
        if not (np.array_equal(d["level"], first["level"]) and np.array_equal(d["ixyz"], first["ixyz"])):
            raise ValueError(f"{path}: not on the grid of {paths[0]}")
        if dim == 3:
            hvy[:, c, g:g + Bs[2], g:g + Bs[1], g:g + Bs[0]] = d["blocks"][:, :Bs[2], :Bs[1], :Bs[0]]
        else:
            hvy[:, c, 0, g:g + Bs[1], g:g + Bs[0]] = d["blocks"][:, :Bs[1], :Bs[0]]
    return {"hvy": hvy, "level": first["level"], "ixyz": first["ixyz"], "treecode": first["treecode"], "time": float(first["attrs"]["time"][0]),
            "iteration": int(first["attrs"]["iteration"][0]), "refinement_status": first["refinement_status"], "attrs": first["attrs"]}
