"""Mesh / time-series I/O around the hot path, host side, no h5py needed.

* ``H5File``       reads the HDF5 subset dolfinx's ``XDMFFile`` writes (``Code/KPP/KPP_exact.py:108-109,165``):
                   version-0 superblock, symbol-table groups, version-1 object headers, contiguous datasets of
                   little-endian integers / IEEE floats.  Enough to load the reference's ``Data/*.h5`` meshes and
                   stored solutions by dataset path.
* ``read_xdmf``    follows an ``.xdmf`` file (``Format="HDF"`` items resolved through ``H5File``, ``Format="Binary"``
                   items through the sidecar file) -> mesh + time series.
* ``XdmfWriter``   the writing side of the time loops (``xdmf.write_mesh`` / ``xdmf.write_function(uh, t)``): an
                   XDMF 3 file whose heavy data is a raw little-endian sidecar (``Format="Binary"`` with ``Seek``),
                   which ParaView opens directly.  Snapshots are appended as they arrive, so a run can stream
                   device->host copies into it.
"""
from __future__ import annotations

import os
import struct
import xml.etree.ElementTree as ET

import numpy as np

_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(ValueError):
    pass


class H5File:
    """Read-only view of a small HDF5 file: ``.datasets`` maps ``/group/name`` -> (shape, dtype, offset)."""

    def __init__(self, path):
        self.path = path
        with open(path, "rb") as f:
            self._b = f.read()
        b = self._b
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise H5Error(f"{path}: not an HDF5 file")
        if b[8] != 0 or b[13] != 8 or b[14] != 8:
            raise H5Error(f"{path}: only version-0 superblocks with 8-byte offsets are supported")
        self._base = struct.unpack_from("<Q", b, 24)[0]
        self.datasets = {}
        # root symbol-table entry at byte 56: name offset, object header, cache type, reserved, scratch
        _, root_hdr, cache, _ = struct.unpack_from("<QQII", b, 56)
        if cache == 1:
            btree, heap = struct.unpack_from("<QQ", b, 80)
            self._walk_group("", btree, heap)
        else:
            self._visit("", root_hdr)

    # -- structure
    def _messages(self, addr):
        b = self._b
        a = self._base + addr
        version, _, nmsg, _, size = struct.unpack_from("<BBHII", b, a)
        if version != 1:
            raise H5Error(f"{self.path}: object header version {version} not supported")
        blocks = [(a + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, p)
                body = p + 8
                if mtype == 0x0010:  # continuation
                    off, ln = struct.unpack_from("<QQ", b, body)
                    blocks.append((self._base + off, ln))
                out.append((mtype, body, msize))
                p = body + msize
        return out

    def _visit(self, name, hdr_addr):
        b = self._b
        shape = dtype = layout = None
        for mtype, p, size in self._messages(hdr_addr):
            if mtype == 0x0011:  # symbol table: this object is a group
                btree, heap = struct.unpack_from("<QQ", b, p)
                self._walk_group(name, btree, heap)
                return
            if mtype == 0x0001:  # dataspace
                ver, rank, flags = struct.unpack_from("<BBB", b, p)
                q = p + (8 if ver == 1 else 4)
                shape = struct.unpack_from(f"<{rank}Q", b, q)
            elif mtype == 0x0003:  # datatype
                cv, bits0 = struct.unpack_from("<BB", b, p)
                cls, nbytes = cv & 0x0F, struct.unpack_from("<I", b, p + 4)[0]
                if bits0 & 1:
                    raise H5Error(f"{self.path}:{name}: big-endian data not supported")
                if cls == 0:
                    dtype = np.dtype(f"<{'i' if bits0 & 8 else 'u'}{nbytes}")
                elif cls == 1:
                    dtype = np.dtype(f"<f{nbytes}")
            elif mtype == 0x0008:  # data layout
                ver, lclass = struct.unpack_from("<BB", b, p)
                if ver == 3 and lclass == 1:
                    layout = struct.unpack_from("<QQ", b, p + 2)
                elif ver == 3 and lclass == 0:
                    layout = ("compact", p + 4, struct.unpack_from("<H", b, p + 2)[0])
                else:
                    layout = ("unsupported", ver, lclass)
        if shape is not None and dtype is not None and layout is not None:
            self.datasets[name] = (tuple(int(s) for s in shape), dtype, layout)

    def _walk_group(self, prefix, btree_addr, heap_addr):
        b = self._b
        h = self._base + heap_addr
        if b[h:h + 4] != b"HEAP":
            raise H5Error(f"{self.path}: bad local heap")
        data = self._base + struct.unpack_from("<Q", b, h + 24)[0]

        def node(addr):
            a = self._base + addr
            sig = b[a:a + 4]
            if sig == b"TREE":
                _ntype, _level, used = struct.unpack_from("<BBH", b, a + 4)
                p = a + 24
                for k in range(used):
                    child = struct.unpack_from("<Q", b, p + 8 + 16 * k)[0]
                    node(child)
            elif sig == b"SNOD":
                nsym = struct.unpack_from("<H", b, a + 6)[0]
                for k in range(nsym):
                    e = a + 8 + 40 * k
                    name_off, hdr = struct.unpack_from("<QQ", b, e)
                    s = data + name_off
                    name = b[s:b.index(b"\0", s)].decode()
                    self._visit(f"{prefix}/{name}", hdr)
            else:
                raise H5Error(f"{self.path}: unexpected node signature {sig!r}")

        node(btree_addr)

    # -- data
    def read(self, name):
        if name not in self.datasets:
            raise KeyError(f"{self.path}: no dataset {name!r}")
        shape, dtype, layout = self.datasets[name]
        count = int(np.prod(shape)) if shape else 1
        if layout[0] == "compact":
            off = layout[1]
        elif layout[0] == "unsupported":
            raise H5Error(f"{self.path}:{name}: data layout version {layout[1]} class {layout[2]} not supported")
        else:
            if layout[0] == _UNDEF:
                return np.zeros(shape, dtype=dtype)
            off = self._base + layout[0]
        return np.frombuffer(self._b, dtype=dtype, count=count, offset=off).reshape(shape).copy()


def _read_item(item, folder, h5cache):
    dims = tuple(int(d) for d in item.attrib["Dimensions"].split())
    fmt = item.attrib.get("Format", "XML")
    text = (item.text or "").strip()
    if fmt == "HDF":
        fname, dset = text.split(":", 1)
        path = os.path.join(folder, fname)
        if path not in h5cache:
            h5cache[path] = H5File(path)
        return h5cache[path].read(dset).reshape(dims)
    if fmt == "Binary":
        kind = item.attrib.get("NumberType", item.attrib.get("DataType", "Float"))
        prec = int(item.attrib.get("Precision", "8" if kind == "Float" else "4"))
        dt = np.dtype(("<f" if kind == "Float" else "<i") + str(prec))
        return np.fromfile(os.path.join(folder, text), dtype=dt, count=int(np.prod(dims)),
                           offset=int(item.attrib.get("Seek", "0"))).reshape(dims)
    return np.array(text.split(), dtype=np.float64).reshape(dims)


def read_xdmf(path):
    """-> dict(x=(Nn,2|3) f64, cells=(Nc,3) i32, series={name: (times (T,), frames (T, Nn[, k]))})."""
    root = ET.parse(path).getroot()
    folder = os.path.dirname(os.path.abspath(path))
    cache = {}
    dom = root.find("Domain")
    mesh = next(g for g in dom.findall("Grid") if g.attrib.get("GridType", "Uniform") == "Uniform")
    cells = _read_item(mesh.find("Topology/DataItem"), folder, cache).astype(np.int32)
    x = np.asarray(_read_item(mesh.find("Geometry/DataItem"), folder, cache), dtype=np.float64)
    series = {}
    for coll in dom.findall("Grid"):
        if coll.attrib.get("GridType") != "Collection":
            continue
        for g in coll.findall("Grid"):
            t = float(g.find("Time").attrib["Value"])
            for a in g.findall("Attribute"):
                v = _read_item(a.find("DataItem"), folder, cache)
                if v.ndim == 2 and v.shape[1] == 1:
                    v = v[:, 0]
                ts, fs = series.setdefault(a.attrib["Name"], ([], []))
                ts.append(t)
                fs.append(v)
    return {"x": x, "cells": cells, "series": {k: (np.array(ts), np.array(fs)) for k, (ts, fs) in series.items()}}


def read_mesh(path):
    """``(x, cells)`` of an ``.xdmf`` file or of a dolfinx ``.h5`` file (``/Mesh/<name>/{geometry,topology}``)."""
    if path.endswith(".xdmf"):
        d = read_xdmf(path)
        return d["x"], d["cells"]
    f = H5File(path)
    geo = next(k for k in f.datasets if k.startswith("/Mesh/") and k.endswith("/geometry"))
    return np.asarray(f.read(geo), dtype=np.float64), f.read(geo[:-len("geometry")] + "topology").astype(np.int32)


class XdmfWriter:
    """``io.XDMFFile(comm, path, "w")`` + ``write_mesh`` + ``write_function(f, t)`` of the reference loops."""

    def __init__(self, path, x, cells, heavy="binary", flush_every=16):
        """``heavy="binary"``: raw little-endian sidecar ``<name>.bin``; ``heavy="hdf5"``: ``<name>.h5`` laid out
        like dolfinx's (``/Mesh/mesh/{topology,geometry}``, ``/Function/<name>/<t with . -> _>``).
        ``flush_every``: the XML (and, for hdf5, the file's metadata tree, which is re-appended whole on every flush)
        is brought up to date every that many frames and on ``close()`` -- per-frame flushes made both grow
        quadratically with the number of frames."""
        self.path = path
        self._flush_every = max(1, int(flush_every))
        self._unflushed = 0
        folder = os.path.dirname(os.path.abspath(path))
        stem = os.path.splitext(os.path.basename(path))[0]
        x = np.ascontiguousarray(x, dtype="<f8")
        self.nn, self.gdim, self.nc = x.shape[0], x.shape[1], np.asarray(cells).shape[0]
        self._frames = []   # (name, t, offset | dataset path, ncomp)
        self._h5 = self._bin = None
        if heavy == "hdf5":
            self.bin_name = stem + ".h5"
            self._h5 = H5Writer(os.path.join(folder, self.bin_name))
            self._h5.write("/Mesh/mesh/topology", np.ascontiguousarray(cells, dtype="<i8"))
            self._h5.write("/Mesh/mesh/geometry", x)
            self._h5.flush()
        elif heavy == "binary":
            self.bin_name = stem + ".bin"
            self._bin = open(os.path.join(folder, self.bin_name), "wb")
            self._topo_off = self._append(np.ascontiguousarray(cells, dtype="<i4"))
            self._geom_off = self._append(x)
        else:
            raise ValueError("heavy must be 'binary' or 'hdf5'")
        self._flush_xml()

    def _append(self, a):
        off = self._bin.tell()
        self._bin.write(a.tobytes())
        return off

    def write_function(self, f, t, name=None):
        a = f.x.array if hasattr(f, "x") and hasattr(f.x, "array") else f
        a = np.ascontiguousarray(a, dtype="<f8").reshape(self.nn, -1)
        name = name or getattr(f, "name", "f")
        if self._h5:
            # dolfinx names the dataset boost::lexical_cast<std::string>(t) with '.' -> '_': 17 significant digits,
            # no trailing zeros ("1", "0_01", "0_30000000000000004"), i.e. C's %.17g -- not Python's repr ("1.0")
            base = f"/Function/{name}/{format(float(t), '.17g').replace('.', '_')}"
            where, k = base, 0
            taken = {w for (_, _, w, _) in self._frames}
            while where in taken:      # the same (name, t) written twice: keep both frames
                k += 1
                where = f"{base}_{k}"
            self._h5.write(where, a)
        else:
            where = self._append(a)
        self._frames.append((name, float(t), where, a.shape[1]))
        self._unflushed += 1
        if self._unflushed >= self._flush_every:
            self.flush()

    def flush(self):
        """Make the files on disk complete and consistent up to the last frame written."""
        if self._h5:
            self._h5.flush()
        else:
            self._bin.flush()
        self._flush_xml()
        self._unflushed = 0

    def _item(self, dims, kind, prec, where):
        if self._h5:
            return f'<DataItem Dimensions="{dims}" NumberType="{kind}" Precision="{prec}" Format="HDF">{self.bin_name}:{where}</DataItem>'
        return (f'<DataItem Dimensions="{dims}" NumberType="{kind}" Precision="{prec}" Format="Binary" Endian="Little" '
                f'Seek="{where}">{self.bin_name}</DataItem>')

    def _flush_xml(self):
        topo = self._item(f"{self.nc} 3", "Int", 8 if self._h5 else 4, "/Mesh/mesh/topology" if self._h5 else self._topo_off)
        geom = self._item(f"{self.nn} {self.gdim}", "Float", 8, "/Mesh/mesh/geometry" if self._h5 else self._geom_off)
        gtype = "XY" if self.gdim == 2 else "XYZ"
        mesh = [f'<Topology TopologyType="Triangle" NumberOfElements="{self.nc}" NodesPerElement="3">', "  " + topo,
                "</Topology>", f'<Geometry GeometryType="{gtype}">', "  " + geom, "</Geometry>"]
        L = ['<?xml version="1.0"?>', '<Xdmf Version="3.0">', "  <Domain>", '    <Grid Name="mesh" GridType="Uniform">']
        L += ["      " + m for m in mesh] + ["    </Grid>"]
        names = []
        for nm, *_ in self._frames:
            if nm not in names:
                names.append(nm)
        for nm in names:
            L.append(f'    <Grid Name="{nm}" GridType="Collection" CollectionType="Temporal">')
            for n2, t, where, k in self._frames:
                if n2 != nm:
                    continue
                L.append(f'      <Grid Name="{nm}" GridType="Uniform">')
                L += ["        " + m for m in mesh]
                L += [f'        <Time Value="{t!r}" />',
                      f'        <Attribute Name="{nm}" AttributeType="{"Scalar" if k == 1 else "Vector"}" Center="Node">',
                      "          " + self._item(f"{self.nn} {k}", "Float", 8, where), "        </Attribute>", "      </Grid>"]
            L.append("    </Grid>")
        L += ["  </Domain>", "</Xdmf>"]
        tmp = self.path + ".tmp"
        with open(tmp, "w") as f:
            f.write("\n".join(L) + "\n")
        os.replace(tmp, self.path)

    def close(self):
        if self._bin or self._h5:
            self.flush()
        if self._bin:
            self._bin.close()
            self._bin = None
        if self._h5:
            self._h5.close()
            self._h5 = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


# --------------------------------------------------------------------------------------- HDF5 writing side
class H5Writer:
    """Writes the same HDF5 subset ``H5File`` reads -- what dolfinx's ``XDMFFile`` produces with the library's
    default ("earliest") file format: version-0 superblock, old-style groups (local heap + version-1 B-tree of
    symbol nodes, leaf K = 4, internal K = 16), version-1 object headers, contiguous little-endian datasets.

    Raw data is appended as ``write`` is called; ``flush`` appends a fresh copy of the (small) metadata tree and
    points the superblock at it, so the file is a complete HDF5 file after every flush and a run can stream
    snapshots into it.  Message encodings are byte-identical to the ones libhdf5 wrote into the reference's
    ``Data/*.h5`` (checked in ``tests/test_io.py`` when the reference tree is mounted); the files have not been
    opened with libhdf5 itself in this image (no h5py / HDF5 tools here)."""

    _SNOD_MAX, _TREE_MAX = 8, 32          # 2 x leaf K, 2 x internal K
    _F64 = bytes.fromhex("11203f000800000000004000340b0034ff03000000000000")
    _FILL = bytes.fromhex("0202020100000000")

    def __init__(self, path):
        self.path = path
        self._f = open(path, "wb+")
        self._f.write(b"\0" * 96)
        self._dsets = {}                    # "/a/b/name" -> (shape, dtype, address, nbytes)
        self._mtime = 0

    # -- raw data
    def _alloc(self, blob):
        f = self._f
        f.seek(0, 2)
        pad = (-f.tell()) % 8
        f.write(b"\0" * pad)
        addr = f.tell()
        f.write(blob)
        return addr

    def write(self, name, array):
        a = np.asarray(array)
        if a.dtype.kind == "f":
            a = np.ascontiguousarray(a, dtype="<f8")
        elif a.dtype.kind in "iu":
            a = np.ascontiguousarray(a, dtype="<i8" if a.dtype.itemsize == 8 else "<i4")
        else:
            raise H5Error(f"unsupported dtype {a.dtype}")
        if not name.startswith("/") or name in self._dsets:
            raise H5Error(f"bad or duplicate dataset path {name!r}")
        self._dsets[name] = (a.shape, a.dtype, self._alloc(a.tobytes()), a.nbytes)

    # -- metadata
    @staticmethod
    def _msg(mtype, body, flags=0):
        assert len(body) % 8 == 0
        return struct.pack("<HHB3x", mtype, len(body), flags) + body

    def _dataset_header(self, shape, dtype, addr, nbytes):
        rank = len(shape)
        space = struct.pack("<BBB5x", 1, rank, 1) + struct.pack(f"<{2 * rank}Q", *shape, *shape)
        if dtype.kind == "f":
            dt = self._F64
        else:
            dt = struct.pack("<BBBBIHH4x", 0x10, 0x08, 0, 0, dtype.itemsize, 0, 8 * dtype.itemsize)
        layout = struct.pack("<BBQQ6x", 3, 1, addr, nbytes)
        msgs = [self._msg(0x0001, space), self._msg(0x0003, dt, 1), self._msg(0x0005, self._FILL, 1),
                self._msg(0x0008, layout), self._msg(0x0012, struct.pack("<B3xI", 1, self._mtime))]
        body = b"".join(msgs)
        return self._alloc(struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body)

    @staticmethod
    def _balanced(items, cap):
        n = len(items)
        k = max(1, -(-n // cap))
        base, extra = divmod(n, k)
        out, p = [], 0
        for i in range(k):
            m = base + (1 if i < extra else 0)
            out.append(items[p:p + m])
            p += m
        return out

    def _group(self, entries):
        """entries: name -> (header address, cache type, scratch 16 bytes).  Returns (btree, heap) addresses."""
        names = sorted(entries)                                   # strcmp order (ASCII names)
        heap, off = bytearray(8), {}
        for nm in names:
            off[nm] = len(heap)
            raw = nm.encode() + b"\0"
            heap += raw + b"\0" * ((-len(raw)) % 8)
        free_at = len(heap)
        heap += struct.pack("<QQ", 1, 16)                          # one free block: (no next, its own size)
        data_addr = self._alloc(bytes(heap))
        heap_addr = self._alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), free_at, data_addr))
        # symbol nodes
        level = []                                                # (address, heap offset of the largest name)
        for chunk in self._balanced(names, self._SNOD_MAX):
            body = bytearray()
            for nm in chunk:
                hdr, ctype, scratch = entries[nm]
                body += struct.pack("<QQII", off[nm], hdr, ctype, 0) + scratch
            body += b"\0" * (40 * (self._SNOD_MAX - len(chunk)))
            level.append((self._alloc(b"SNOD" + struct.pack("<BBH", 1, 0, len(chunk)) + bytes(body)), off[chunk[-1]]))
        # B-tree levels, bottom up; sibling links need the node addresses first
        depth = 0
        while True:
            groups = self._balanced(level, self._TREE_MAX)
            size = 24 + 8 * (self._TREE_MAX + 1) + 8 * self._TREE_MAX   # header + 2K+1 keys + 2K children (_TREE_MAX = 2K)
            addrs = [self._alloc(b"\0" * size) for _ in groups]
            nxt, left_key = [], 0
            for i, g in enumerate(groups):
                node = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, depth, len(g), addrs[i - 1] if i else _UNDEF,
                                                       addrs[i + 1] if i + 1 < len(groups) else _UNDEF))
                node += struct.pack("<Q", left_key)
                for child, kmax in g:
                    node += struct.pack("<QQ", child, kmax)
                node += b"\0" * (size - len(node))
                self._f.seek(addrs[i])
                self._f.write(node)
                left_key = g[-1][1]
                nxt.append((addrs[i], left_key))
            if len(nxt) == 1:
                return nxt[0][0], heap_addr
            level, depth = nxt, depth + 1

    def flush(self):
        import time

        self._mtime = int(time.time())
        tree = {}
        for path, rec in self._dsets.items():
            parts = path.strip("/").split("/")
            d = tree
            for p in parts[:-1]:
                d = d.setdefault(p, {})
                if not isinstance(d, dict):
                    raise H5Error(f"{path}: a dataset is used as a group")
            d[parts[-1]] = rec

        def emit(node):
            ent = {}
            for nm, child in node.items():
                if isinstance(child, dict):
                    bt, hp = emit(child)
                    stab = struct.pack("<QQ", bt, hp)
                    hdr = self._alloc(struct.pack("<BBHII4x", 1, 0, 1, 1, 24) + self._msg(0x0011, stab))
                    ent[nm] = (hdr, 1, stab)
                else:
                    ent[nm] = (self._dataset_header(*child), 0, b"\0" * 16)
            return self._group(ent)

        bt, hp = emit(tree)
        stab = struct.pack("<QQ", bt, hp)
        root = self._alloc(struct.pack("<BBHII4x", 1, 0, 1, 1, 24) + self._msg(0x0011, stab))
        self._f.seek(0, 2)
        eof = self._f.tell()
        sb = (b"\x89HDF\r\n\x1a\n" + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", 4, 16, 0) +
              struct.pack("<QQQQ", 0, _UNDEF, eof, _UNDEF) + struct.pack("<QQII", 0, root, 1, 0) + stab)
        assert len(sb) == 96
        self._f.seek(0)
        self._f.write(sb)
        self._f.flush()

    def close(self):
        if self._f:
            self.flush()
            self._f.close()
            self._f = None
