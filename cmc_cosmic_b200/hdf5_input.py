"""Reader of the reference's initial-condition files (``input.hdf5``; reference src/libs/fitslib.c:209-448).

COSMIC writes them with pandas' fixed format: two groups ``CLUS_OBJ_DATA`` and ``CLUS_BINARY_DATA`` whose dataset ``block0_values`` is a
contiguous little-endian float64 array [N + 2][8] resp. [N_bin + 1][>= 11]; the scalars the reference reads with H5Aread (NOBJ, NBINARY, Z,
RVIR, RTID, MCLUS) are attributes of ``CLUS_OBJ_DATA/block0_values``.  No HDF5 library exists in this image, so this module parses the
part of the file format those files use -- superblock version 0, version-1 object headers with continuation blocks, symbol-table groups
(version-1 B-tree + local heap), simple dataspaces, fixed-point / floating-point datatypes, contiguous (and compact) layouts, attribute
messages of versions 1-3 -- and refuses anything else.  Host-side I/O: nothing here touches the GPU.

Columns (fitslib.c:381-420):  objects  id, k, m, Reff, r, vr, vt, binind;
binaries  index, id1, k1, m1, Reff1, id2, k2, m2, Reff2, a, e  (the first 11 of the row).
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5FormatError(ValueError):
    pass


class _File:
    def __init__(self, raw):
        self.raw = raw
        if raw[:8] != b"\x89HDF\r\n\x1a\n":
            raise Hdf5FormatError("not an HDF5 file")
        if raw[8] != 0:
            raise Hdf5FormatError(f"superblock version {raw[8]} (only version 0 is read)")
        if raw[13] != 8 or raw[14] != 8:
            raise Hdf5FormatError("only 8-byte offsets and lengths are read")
        self.base = struct.unpack_from("<Q", raw, 24)[0]
        self.root = struct.unpack_from("<Q", raw, 56 + 8)[0]      # root symbol-table entry: link name offset, object header address

    # ---- object headers (version 1) ----------------------------------------------------------------------------------------------
    def messages(self, addr):
        raw = self.raw
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", raw, addr)
        if ver != 1:
            raise Hdf5FormatError(f"object header version {ver} at {addr} (only version 1 is read)")
        blocks, out = [(addr + 16, addr + 16 + hsize)], []
        while blocks and len(out) < nmsg:
            p, end = blocks.pop(0)
            while p + 8 <= end and len(out) < nmsg:
                t, sz, _fl = struct.unpack_from("<HHB", raw, p)
                body = raw[p + 8:p + 8 + sz]
                if t == 0x10:                                      # continuation
                    off, ln = struct.unpack_from("<QQ", body, 0)
                    blocks.append((off + self.base, off + self.base + ln))
                out.append((t, body))
                p += 8 + sz
        return out

    # ---- groups: symbol table message -> B-tree of symbol nodes + local heap of names ---------------------------------------------
    def children(self, addr):
        for t, body in self.messages(addr):
            if t == 0x11:
                btree, heap = struct.unpack_from("<QQ", body, 0)
                return dict(self._walk_btree(btree + self.base, self._heap_data(heap + self.base)))
        raise Hdf5FormatError(f"object at {addr} is not an old-style group")

    def _heap_data(self, addr):
        raw = self.raw
        if raw[addr:addr + 4] != b"HEAP":
            raise Hdf5FormatError("local heap signature missing")
        return struct.unpack_from("<Q", raw, addr + 24)[0] + self.base

    def _walk_btree(self, addr, heap_data):
        raw = self.raw
        if raw[addr:addr + 4] != b"TREE":
            raise Hdf5FormatError("B-tree signature missing")
        ntype, level, used = struct.unpack_from("<BBH", raw, addr + 4)
        if ntype != 0:
            raise Hdf5FormatError("not a group B-tree")
        p = addr + 24                                              # keys and children alternate: key0 child0 key1 child1 ... key_n
        for i in range(used):
            child = struct.unpack_from("<Q", raw, p + 8 + 16 * i)[0] + self.base
            if level > 0:
                yield from self._walk_btree(child, heap_data)
            else:
                if raw[child:child + 4] != b"SNOD":
                    raise Hdf5FormatError("symbol node signature missing")
                nsym = struct.unpack_from("<H", raw, child + 6)[0]
                for k in range(nsym):
                    e = child + 8 + 40 * k
                    name_off, ohdr = struct.unpack_from("<QQ", raw, e)
                    s = heap_data + name_off
                    yield raw[s:raw.index(b"\0", s)].decode(), ohdr + self.base

    def open(self, path):
        addr = self.root + self.base
        for part in path.strip("/").split("/"):
            kids = self.children(addr)
            if part not in kids:
                raise KeyError(f"{path}: no '{part}' (has {sorted(kids)})")
            addr = kids[part]
        return addr

    # ---- datasets and attributes -----------------------------------------------------------------------------------------------------
    @staticmethod
    def _dtype(body):
        cls, size = body[0] & 0x0F, struct.unpack_from("<I", body, 4)[0]
        bits0 = body[1]
        if cls in (0, 1) and bits0 & 1:
            raise Hdf5FormatError("big-endian data")
        if cls == 1 and size in (4, 8):
            return np.dtype(f"<f{size}")
        if cls == 0 and size in (1, 2, 4, 8):
            return np.dtype(("<i" if bits0 & 8 else "<u") + str(size))
        return None                                                # strings, variable-length ...: not needed by the reader

    @staticmethod
    def _shape(body):
        ver, rank, flags = body[0], body[1], body[2]
        off = 8 if ver == 1 else 4
        return tuple(struct.unpack_from("<Q", body, off + 8 * i)[0] for i in range(rank))

    def dataset(self, addr):
        dt = shape = data = None
        for t, body in self.messages(addr):
            if t == 0x03:
                dt = self._dtype(body)
            elif t == 0x01:
                shape = self._shape(body)
            elif t == 0x08:
                if body[0] != 3:
                    raise Hdf5FormatError(f"data layout message version {body[0]}")
                if body[1] == 1:                                   # contiguous
                    a, n = struct.unpack_from("<QQ", body, 2)
                    if a == UNDEF:
                        raise Hdf5FormatError("dataset without storage")
                    data = self.raw[a + self.base:a + self.base + n]
                elif body[1] == 0:                                 # compact
                    n = struct.unpack_from("<H", body, 2)[0]
                    data = body[4:4 + n]
                else:
                    raise Hdf5FormatError("chunked datasets are not read (the reference's inputs are written contiguous)")
        if dt is None or shape is None or data is None:
            raise Hdf5FormatError("incomplete dataset header")
        return np.frombuffer(data, dt, int(np.prod(shape, dtype=np.int64))).reshape(shape).copy()

    def attributes(self, addr):
        out = {}
        for t, body in self.messages(addr):
            if t != 0x0C:
                continue
            ver = body[0]
            nsz, tsz, ssz = struct.unpack_from("<HHH", body, 2)
            p = 8 if ver < 3 else 9
            pad = (lambda n: (n + 7) & ~7) if ver == 1 else (lambda n: n)
            name = body[p:p + nsz].split(b"\0")[0].decode()
            p += pad(nsz)
            dt = self._dtype(body[p:p + tsz])
            p += pad(tsz)
            shape = self._shape(body[p:p + ssz]) if ssz >= 4 and body[p + 1] > 0 else ()
            p += pad(ssz)
            if dt is None:
                continue
            n = int(np.prod(shape, dtype=np.int64)) if shape else 1
            val = np.frombuffer(body[p:p + n * dt.itemsize], dt, n)
            out[name] = val.reshape(shape) if shape else val[0]
        return out


OBJ_COLUMNS = ("id", "k", "m", "Reff", "r", "vr", "vt", "binind")
BIN_COLUMNS = ("index", "id1", "k1", "m1", "Reff1", "id2", "k2", "m2", "Reff2", "a", "e")


def read_input(path):
    """cmc_read_hdf5_file (reference src/libs/fitslib.c:209-448): returns dict(NOBJ, NBINARY, Z, RVIR, RTID, MCLUS, obj = {column: array
    [NOBJ + 2]}, bin = {column: array [NBINARY + 1]}); row 0 of either table and row NOBJ + 1 of the objects are the reference's sentinels."""
    f = _File(open(path, "rb").read())
    oaddr = f.open("CLUS_OBJ_DATA/block0_values")
    baddr = f.open("CLUS_BINARY_DATA/block0_values")
    attrs = f.attributes(oaddr)
    missing = [k for k in ("NOBJ", "NBINARY", "Z", "RVIR", "RTID", "MCLUS") if k not in attrs]
    if missing:
        raise Hdf5FormatError(f"attributes {missing} missing on CLUS_OBJ_DATA/block0_values")
    obj, binr = f.dataset(oaddr), f.dataset(baddr)
    nobj, nbin = int(attrs["NOBJ"]), int(attrs["NBINARY"])
    if obj.ndim != 2 or obj.shape[1] < 8 or obj.shape[0] < nobj + 2:
        raise Hdf5FormatError(f"CLUS_OBJ_DATA/block0_values has shape {obj.shape}, NOBJ = {nobj}")
    if binr.ndim != 2 or binr.shape[1] < 11 or binr.shape[0] < nbin + 1:
        raise Hdf5FormatError(f"CLUS_BINARY_DATA/block0_values has shape {binr.shape}, NBINARY = {nbin}")
    out = {k: (int(attrs[k]) if k in ("NOBJ", "NBINARY") else float(attrs[k])) for k in ("NOBJ", "NBINARY", "Z", "RVIR", "RTID", "MCLUS")}
    out["obj"] = {c: obj[:nobj + 2, i].copy() for i, c in enumerate(OBJ_COLUMNS)}
    out["bin"] = {c: binr[:nbin + 1, i].copy() for i, c in enumerate(BIN_COLUMNS)}
    return out
