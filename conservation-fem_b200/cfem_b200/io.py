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

    def __init__(self, path, x, cells):
        self.path = path
        self.bin_name = os.path.splitext(os.path.basename(path))[0] + ".bin"
        self._bin = open(os.path.join(os.path.dirname(os.path.abspath(path)), self.bin_name), "wb")
        x = np.ascontiguousarray(x, dtype="<f8")
        cells = np.ascontiguousarray(cells, dtype="<i4")
        self.nn, self.gdim, self.nc = x.shape[0], x.shape[1], cells.shape[0]
        self._topo_off = self._append(cells)
        self._geom_off = self._append(x)
        self._frames = []   # (name, t, offset, ncomp)
        self._flush_xml()

    def _append(self, a):
        off = self._bin.tell()
        self._bin.write(a.tobytes())
        return off

    def write_function(self, f, t, name=None):
        a = f.x.array if hasattr(f, "x") and hasattr(f.x, "array") else f
        a = np.ascontiguousarray(a, dtype="<f8").reshape(self.nn, -1)
        self._frames.append((name or getattr(f, "name", "f"), float(t), self._append(a), a.shape[1]))
        self._bin.flush()
        self._flush_xml()

    def _flush_xml(self):
        L = ['<?xml version="1.0"?>', '<Xdmf Version="3.0">', "  <Domain>", '    <Grid Name="mesh" GridType="Uniform">',
             f'      <Topology TopologyType="Triangle" NumberOfElements="{self.nc}" NodesPerElement="3">',
             f'        <DataItem Dimensions="{self.nc} 3" NumberType="Int" Precision="4" Format="Binary" Endian="Little" Seek="{self._topo_off}">{self.bin_name}</DataItem>',
             "      </Topology>", f'      <Geometry GeometryType="{"XY" if self.gdim == 2 else "XYZ"}">',
             f'        <DataItem Dimensions="{self.nn} {self.gdim}" NumberType="Float" Precision="8" Format="Binary" Endian="Little" Seek="{self._geom_off}">{self.bin_name}</DataItem>',
             "      </Geometry>", "    </Grid>"]
        names = []
        for nm, *_ in self._frames:
            if nm not in names:
                names.append(nm)
        for nm in names:
            L.append(f'    <Grid Name="{nm}" GridType="Collection" CollectionType="Temporal">')
            for n2, t, off, k in self._frames:
                if n2 != nm:
                    continue
                L += [f'      <Grid Name="{nm}" GridType="Uniform">',
                      f'        <Topology TopologyType="Triangle" NumberOfElements="{self.nc}" NodesPerElement="3">',
                      f'          <DataItem Dimensions="{self.nc} 3" NumberType="Int" Precision="4" Format="Binary" Endian="Little" Seek="{self._topo_off}">{self.bin_name}</DataItem>',
                      "        </Topology>", f'        <Geometry GeometryType="{"XY" if self.gdim == 2 else "XYZ"}">',
                      f'          <DataItem Dimensions="{self.nn} {self.gdim}" NumberType="Float" Precision="8" Format="Binary" Endian="Little" Seek="{self._geom_off}">{self.bin_name}</DataItem>',
                      "        </Geometry>", f'        <Time Value="{t!r}" />',
                      f'        <Attribute Name="{nm}" AttributeType="{"Scalar" if k == 1 else "Vector"}" Center="Node">',
                      f'          <DataItem Dimensions="{self.nn} {k}" NumberType="Float" Precision="8" Format="Binary" Endian="Little" Seek="{off}">{self.bin_name}</DataItem>',
                      "        </Attribute>", "      </Grid>"]
            L.append("    </Grid>")
        L += ["  </Domain>", "</Xdmf>"]
        tmp = self.path + ".tmp"
        with open(tmp, "w") as f:
            f.write("\n".join(L) + "\n")
        os.replace(tmp, self.path)

    def close(self):
        if self._bin:
            self._bin.close()
            self._bin = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
