"""Minimal reader for the classic-HDF5 files WABBIT writes (TEST INFRASTRUCTURE ONLY).

This module is part of ``oracle/``: it may be imported by ``tests/`` and by the
fixture generators under ``tests/golden/`` only, never by the product path.

WABBIT's ``saveHDF5_tree`` (reference: LIB/MESH/InputOutput.f90:1-280) writes
superblock-v0 files with chunked, unfiltered little-endian datasets
(``blocks``, ``block_treecode_num``, ``level``, ``coords_origin``,
``coords_spacing``, ``refinement_status``, ``lgt_ids``, ``procs``) and scalar /
small-vector attributes on ``blocks``.  h5py is not available in this image, so
this walks the on-disk structures directly.
"""
from __future__ import annotations

import struct
from typing import Dict, Tuple

import numpy as np


class H5Lite:
    def __init__(self, path: str):
        with open(path, "rb") as f:
            self.b = f.read()
        b = self.b
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise ValueError(f"{path}: not an HDF5 file")
        if b[8] != 0:
            raise ValueError("only superblock v0 supported")
        # root symbol table entry starts at byte 56: link name off(8) objhdr(8) cache(4) res(4) scratch(16)
        self.root_objhdr = struct.unpack_from("<Q", b, 56 + 8)[0]
        btree, heap = struct.unpack_from("<QQ", b, 56 + 24)
        self.datasets: Dict[str, int] = {}
        self._walk_group(btree, heap)

    # ------------------------------------------------------------------ groups
    def _heap_data(self, heap_addr: int) -> int:
        b = self.b
        assert b[heap_addr:heap_addr + 4] == b"HEAP"
        return struct.unpack_from("<Q", b, heap_addr + 24)[0]

    def _walk_group(self, btree: int, heap: int):
        data_seg = self._heap_data(heap)
        self._walk_gnode(btree, data_seg)

    def _walk_gnode(self, addr: int, data_seg: int):
        b = self.b
        sig = b[addr:addr + 4]
        if sig == b"TREE":
            ntype, level, nent = struct.unpack_from("<BBH", b, addr + 4)
            assert ntype == 0
            p = addr + 8 + 16  # skip siblings
            # keys and children interleaved: key0 child0 key1 child1 ... keyN
            for i in range(nent):
                child = struct.unpack_from("<Q", b, p + 8 + i * 16)[0]
                self._walk_gnode(child, data_seg)
        elif sig == b"SNOD":
            nsym = struct.unpack_from("<H", b, addr + 6)[0]
            p = addr + 8
            for i in range(nsym):
                name_off, objhdr = struct.unpack_from("<QQ", b, p + i * 40)
                s = data_seg + name_off
                e = b.index(b"\x00", s)
                self.datasets[b[s:e].decode()] = objhdr
        else:
            raise ValueError(f"unexpected group node signature {sig!r}")

    # ---------------------------------------------------------- object headers
    def _messages(self, objhdr: int):
        b = self.b
        ver, _, nmsg = struct.unpack_from("<BBH", b, objhdr)
        assert ver == 1
        hdr_size = struct.unpack_from("<I", b, objhdr + 8)[0]
        blocks = [(objhdr + 16, hdr_size)]
        out = []
        while blocks and len(out) < nmsg:
            p, n = blocks.pop(0)
            end = p + n
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, p)
                body = p + 8
                if mtype == 0x10:
                    caddr, clen = struct.unpack_from("<QQ", b, body)
                    blocks.append((caddr, clen))
                out.append((mtype, body, msize))
                p = body + msize
        return out

    @staticmethod
    def _dtype(b: bytes, p: int) -> np.dtype:
        cls = b[p] & 0x0F
        size = struct.unpack_from("<I", b, p + 4)[0]
        if cls == 0:
            signed = (b[p + 1] >> 3) & 1
            return np.dtype(f"<{'i' if signed else 'u'}{size}")
        if cls == 1:
            return np.dtype(f"<f{size}")
        if cls == 3:  # string
            return np.dtype(f"S{size}")
        raise ValueError(f"datatype class {cls} unsupported")

    def _dataspace(self, p: int) -> Tuple[int, ...]:
        b = self.b
        ver, rank, flags = struct.unpack_from("<BBB", b, p)
        off = 8 if ver == 1 else 4
        return tuple(struct.unpack_from("<Q", b, p + off + 8 * i)[0] for i in range(rank))

    # ---------------------------------------------------------------- datasets
    def read(self, name: str) -> np.ndarray:
        b = self.b
        shape = dt = layout = None
        for mtype, p, _n in self._messages(self.datasets[name]):
            if mtype == 0x01:
                shape = self._dataspace(p)
            elif mtype == 0x03:
                dt = self._dtype(b, p)
            elif mtype == 0x08:
                layout = p
        ver, cls = b[layout], b[layout + 1]
        assert ver == 3
        n = int(np.prod(shape)) if shape else 1
        if cls == 1:  # contiguous
            addr, size = struct.unpack_from("<QQ", b, layout + 2)
            return np.frombuffer(b, dt, n, addr).reshape(shape).copy()
        if cls == 0:  # compact
            size = struct.unpack_from("<H", b, layout + 2)[0]
            return np.frombuffer(b, dt, n, layout + 4).reshape(shape).copy()
        assert cls == 2
        rank1 = b[layout + 2]
        btree = struct.unpack_from("<Q", b, layout + 3)[0]
        cdims = struct.unpack_from(f"<{rank1}I", b, layout + 11)
        out = np.zeros(shape, dt)
        self._read_chunks(btree, rank1, cdims[:-1], out)
        return out

    def _read_chunks(self, addr: int, rank1: int, cdims, out: np.ndarray):
        b = self.b
        assert b[addr:addr + 4] == b"TREE"
        ntype, level, nent = struct.unpack_from("<BBH", b, addr + 4)
        assert ntype == 1
        keysz = 8 + 8 * rank1
        p = addr + 8 + 16
        for i in range(nent):
            kp = p + i * (keysz + 8)
            nbytes, _mask = struct.unpack_from("<II", b, kp)
            offs = struct.unpack_from(f"<{rank1}Q", b, kp + 8)[:-1]
            child = struct.unpack_from("<Q", b, kp + keysz)[0]
            if level > 0:
                self._read_chunks(child, rank1, cdims, out)
                continue
            chunk = np.frombuffer(b, out.dtype, nbytes // out.dtype.itemsize, child).reshape(cdims)
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, out.shape))
            csl = tuple(slice(0, s.stop - s.start) for s in sl)
            out[sl] = chunk[csl]

    def attrs(self, name: str) -> Dict[str, np.ndarray]:
        b = self.b
        res = {}
        for mtype, p, _n in self._messages(self.datasets[name]):
            if mtype != 0x0C:
                continue
            ver = b[p]
            nsz, dsz, ssz = struct.unpack_from("<HHH", b, p + 2)
            assert ver == 1
            pad = lambda x: (x + 7) & ~7
            q = p + 8
            aname = b[q:q + nsz].split(b"\x00")[0].decode()
            q += pad(nsz)
            dt = self._dtype(b, q)
            q += pad(dsz)
            shape = self._dataspace(q) if ssz >= 8 else ()
            q += pad(ssz)
            n = int(np.prod(shape)) if shape else 1
            res[aname] = np.frombuffer(b, dt, n, q).copy()
        return res


def read_wabbit(path: str):
    """Return dict with blocks (Nb, [z,] y, x), treecode, level, origin, spacing and attrs."""
    f = H5Lite(path)
    d = {
        "blocks": f.read("blocks"),
        "attrs": f.attrs("blocks"),
    }
    for k_out, k_in in (("treecode", "block_treecode_num"), ("level", "level"),
                        ("origin", "coords_origin"), ("spacing", "coords_spacing"),
                        ("refinement_status", "refinement_status")):
        if k_in in f.datasets:
            d[k_out] = f.read(k_in)
    return d
