"""Minimal zarr-v2 array I/O with the attrs funlib.persistence writes (offset, voxel_size, axis_names,
units, types) — `zarr` / `numcodecs` / `funlib.persistence` are not available in this image.

Supports C-order arrays, '.' or '/' chunk keys, compressor null, zlib, or blosc (the zarr / funlib default, what
`bs predict` writes): blosc frames are parsed in bootstrapper_b200/blosc1.py with LZ4 / Zstandard payloads decoded through
pyarrow and zlib through the standard library; bit-shuffle and BloscLZ / Snappy payloads raise.
"""
import json
import os
import zlib

import numpy as np


class ZarrArray:
    def __init__(self, path, mode="r"):
        self.path = path
        with open(os.path.join(path, ".zarray")) as f:
            meta = json.load(f)
        if meta.get("zarr_format") != 2 or meta.get("order", "C") != "C" or meta.get("filters"):
            raise ValueError(f"{path}: only zarr v2, C order, no filters is supported")
        comp = meta.get("compressor")
        if comp is not None and comp.get("id") not in ("zlib", "blosc"):
            raise ValueError(f"{path}: compressor {comp.get('id')!r} is not supported (null, zlib or blosc)")
        if comp is not None and comp.get("id") == "blosc" and (comp.get("shuffle", 1) == 2 or
                                                               (comp.get("shuffle", 1) == -1 and np.dtype(meta["dtype"]).itemsize == 1)):
            raise ValueError(f"{path}: blosc bit-shuffle is not supported")
        self.compressor = comp
        self.shape = tuple(meta["shape"])
        self.chunks = tuple(meta["chunks"])
        self.dtype = np.dtype(meta["dtype"])
        self.fill_value = meta.get("fill_value") or 0
        self.sep = meta.get("dimension_separator", ".")
        self.attrs = {}
        ap = os.path.join(path, ".zattrs")
        if os.path.exists(ap):
            with open(ap) as f:
                self.attrs = json.load(f)
        nsp = len(self.attrs.get("voxel_size", self.shape[-3:]))
        self.voxel_size = tuple(int(v) for v in self.attrs.get("voxel_size", (1,) * nsp))
        self.offset = tuple(int(v) for v in self.attrs.get("offset", (0,) * nsp))
        self.axis_names = self.attrs.get("axis_names")
        self.units = self.attrs.get("units")
        self.types = self.attrs.get("types")
        self.mode = mode

    # ---- geometry in world units (spatial dims are the trailing ones)
    @property
    def spatial_shape(self):
        return self.shape[-len(self.voxel_size):]

    @property
    def roi(self):
        return self.offset, tuple(s * v for s, v in zip(self.spatial_shape, self.voxel_size))

    @property
    def chunk_shape(self):
        return self.chunks

    def _chunk_path(self, idx):
        return os.path.join(self.path, self.sep.join(str(i) for i in idx))

    def _read_chunk(self, idx):
        p = self._chunk_path(idx)
        if not os.path.exists(p):
            return np.full(self.chunks, self.fill_value, dtype=self.dtype)
        with open(p, "rb") as f:
            raw = f.read()
        if self.compressor is not None:
            if self.compressor["id"] == "blosc":
                from . import blosc1
                raw = blosc1.decode(raw)
            else:
                raw = zlib.decompress(raw)
        return np.frombuffer(raw, dtype=self.dtype).reshape(self.chunks)

    def _write_chunk(self, idx, data):
        p = self._chunk_path(idx)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        raw = np.ascontiguousarray(data, dtype=self.dtype).tobytes()
        if self.compressor is not None:
            if self.compressor["id"] == "blosc":
                from . import blosc1
                raw = blosc1.encode(raw, typesize=self.dtype.itemsize, cname=self.compressor.get("cname", "lz4"),
                                    shuffle=1 if self.compressor.get("shuffle", 1) in (1, -1) else 0,
                                    blocksize=self.compressor.get("blocksize", 0))
            else:
                raw = zlib.compress(raw, self.compressor.get("level", 1))
        with open(p, "wb") as f:
            f.write(raw)

    def read(self, start=None, stop=None):
        """array[start:stop] in voxel indices over all dims (must lie inside the array)."""
        start = tuple(start) if start is not None else (0,) * len(self.shape)
        stop = tuple(stop) if stop is not None else self.shape
        out = np.empty(tuple(b - a for a, b in zip(start, stop)), dtype=self.dtype)
        lo = [a // c for a, c in zip(start, self.chunks)]
        hi = [(b - 1) // c + 1 for b, c in zip(stop, self.chunks)]
        for idx in np.ndindex(*[h - l for l, h in zip(lo, hi)]):
            cidx = tuple(i + l for i, l in zip(idx, lo))
            c0 = [i * c for i, c in zip(cidx, self.chunks)]
            src, dst = [], []
            for d in range(len(self.shape)):
                a, b = max(start[d], c0[d]), min(stop[d], c0[d] + self.chunks[d])
                src.append(slice(a - c0[d], b - c0[d]))
                dst.append(slice(a - start[d], b - start[d]))
            out[tuple(dst)] = self._read_chunk(cidx)[tuple(src)]
        return out

    def to_ndarray(self, start, shape, fill_value=0):
        """funlib.persistence Array.to_ndarray(roi, fill_value): the spatial window [start, start + shape) in voxel
        indices of this array (may reach outside it), all leading (channel) dims, filled with fill_value outside."""
        nsp = len(self.voxel_size)
        lead = self.shape[:-nsp]
        out = np.full(tuple(lead) + tuple(int(v) for v in shape), fill_value, dtype=self.dtype)
        lo = [max(int(a), 0) for a in start]
        hi = [min(int(a) + int(n), s) for a, n, s in zip(start, shape, self.spatial_shape)]
        if any(h <= l for l, h in zip(lo, hi)):
            return out
        data = self.read(tuple([0] * len(lead)) + tuple(lo), tuple(lead) + tuple(hi))
        dst = tuple(slice(l - int(a), h - int(a)) for l, h, a in zip(lo, hi, start))
        out[(Ellipsis,) + dst] = data
        return out

    def write(self, data, start=None):
        start = tuple(start) if start is not None else (0,) * len(self.shape)
        stop = tuple(a + s for a, s in zip(start, data.shape))
        lo = [a // c for a, c in zip(start, self.chunks)]
        hi = [(b - 1) // c + 1 for b, c in zip(stop, self.chunks)]
        for idx in np.ndindex(*[h - l for l, h in zip(lo, hi)]):
            cidx = tuple(i + l for i, l in zip(idx, lo))
            c0 = [i * c for i, c in zip(cidx, self.chunks)]
            src, dst = [], []
            full = True
            for d in range(len(self.shape)):
                a, b = max(start[d], c0[d]), min(stop[d], c0[d] + self.chunks[d])
                dst.append(slice(a - c0[d], b - c0[d]))
                src.append(slice(a - start[d], b - start[d]))
                full &= (b - a) == self.chunks[d]
            chunk = np.array(self._read_chunk(cidx)) if not full else np.empty(self.chunks, self.dtype)
            chunk[tuple(dst)] = data[tuple(src)]
            self._write_chunk(cidx, chunk)


def open_ds(path, mode="r"):
    return ZarrArray(path, mode)


def _ensure_groups(path):
    """.zgroup markers from the enclosing *.zarr container down to the array's parent."""
    parts = os.path.abspath(path).split(os.sep)
    roots = [i for i, p in enumerate(parts) if p.endswith(".zarr")]
    if not roots:
        return
    for i in range(roots[0], len(parts) - 1):
        g = os.sep.join(parts[: i + 1])
        os.makedirs(g, exist_ok=True)
        gp = os.path.join(g, ".zgroup")
        if not os.path.exists(gp) and not os.path.exists(os.path.join(g, ".zarray")):
            with open(gp, "w") as f:
                json.dump({"zarr_format": 2}, f)


def prepare_ds(path, shape, offset, voxel_size, dtype, chunk_shape=None, axis_names=None, units=None, types=None,
               compressor=None):
    """Create (overwrite) a zarr-v2 array with funlib.persistence-style attrs."""
    import shutil
    if os.path.exists(path):
        shutil.rmtree(path)
    _ensure_groups(path)
    os.makedirs(path)
    chunk_shape = tuple(chunk_shape) if chunk_shape is not None else tuple(min(s, 256) for s in shape)
    meta = dict(zarr_format=2, shape=list(shape), chunks=list(chunk_shape), dtype=np.dtype(dtype).str,
                compressor=compressor, fill_value=0, order="C", filters=None, dimension_separator=".")
    with open(os.path.join(path, ".zarray"), "w") as f:
        json.dump(meta, f, indent=4)
    nsp = len(voxel_size)
    attrs = dict(offset=[int(v) for v in offset], voxel_size=[int(v) for v in voxel_size],
                 axis_names=list(axis_names) if axis_names else ["c^"] * (len(shape) - nsp) + ["z", "y", "x"][-nsp:],
                 units=list(units) if units else ["nm"] * nsp)
    if types:
        attrs["types"] = list(types)
    with open(os.path.join(path, ".zattrs"), "w") as f:
        json.dump(attrs, f, indent=4)
    return ZarrArray(path, "r+")
