"""ctypes bindings to the oracle's C/C++ restatements (oracle/csrc)."""
import ctypes as C

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        _lib.sk_watershed.restype = C.c_int64
        _lib.sk_label.restype = C.c_int64
        _lib.wz_create.restype = C.c_void_p
        _lib.wz_create_q.restype = C.c_void_p
        _lib.wz_num_edges.restype = C.c_int64
        _lib.wz_num_nodes.restype = C.c_int64
        _lib.wz_merge_until.restype = C.c_int64
        _lib.wz_region_graph.restype = C.c_int64
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def mws_agglom(affs, offsets, strides=None, zero_is_repulsive=True, counters=None):
    """mwatershed.agglom(affs float64 (C, Z, Y, X), offsets, strides) restated (post/mws.py:52-57); declared tie rule D4.
    Returns uint64 labels (1 + smallest raveled voxel index of the cluster)."""
    affs = np.ascontiguousarray(affs, dtype=np.float64)
    assert affs.ndim == 4
    Cn = affs.shape[0]
    off = np.ascontiguousarray(np.asarray(offsets, dtype=np.int64).reshape(Cn, 3))
    st = None if strides is None else np.ascontiguousarray(np.asarray(strides, dtype=np.int64).reshape(Cn, 3))
    shape = np.asarray(affs.shape[1:], dtype=np.int64)
    out = np.empty(affs.shape[1:], dtype=np.uint64)
    cnt = np.zeros(4, dtype=np.int64)
    lib().mws_agglom.restype = C.c_int64
    rc = lib().mws_agglom(_p(affs), C.c_int(Cn), _p(shape), _p(off), _p(st) if st is not None else None,
                          C.c_int(1 if zero_is_repulsive else 0), _p(out), _p(cnt))
    assert rc == 0, "volume too large for the oracle's 32-bit voxel indices"
    if counters is not None:
        counters.update(edges=int(cnt[0]), merges=int(cnt[1]), mutexes=int(cnt[2]), blocked=int(cnt[3]))
    return out


def sk_watershed(image, markers, mask, seed_tie="heap"):
    """skimage.segmentation.watershed(image, markers, mask=mask) restated (post/ws.py:26-28)."""
    image = np.ascontiguousarray(image, dtype=np.float64)
    markers = np.ascontiguousarray(markers, dtype=np.int64)
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    out = np.empty(image.shape, dtype=np.int64)
    shape = np.asarray(image.shape, dtype=np.int64)
    lib().sk_watershed(_p(image), _p(markers), _p(mask), C.c_int(image.ndim), _p(shape), _p(out),
                       C.c_int({"heap": 0, "index": 1}[seed_tie]))
    return out


def sk_label(x):
    """skimage.measure.label(x, return_num=True) restated (watershed_frags.py:222)."""
    x = np.ascontiguousarray(x, dtype=np.int64)
    out = np.empty(x.shape, dtype=np.int64)
    shape = np.asarray(x.shape, dtype=np.int64)
    n = lib().sk_label(_p(x), C.c_int(x.ndim), _p(shape), _p(out))
    return out, int(n)


class Waterz:
    """State of one waterz.agglomerate call (restated)."""

    def __init__(self, affs, frags, queue_bins, stats_mode="faithful", keep_cheaper=True, quantile=0, initmax=False):
        assert affs.ndim == 4 and affs.shape[0] == 3
        if affs.dtype == np.uint8:
            dt = 0
        else:
            affs = affs.astype(np.float32, copy=False)
            dt = 1
        self.affs = np.ascontiguousarray(affs)
        self.frags = np.ascontiguousarray(frags, dtype=np.uint64)
        Z, Y, X = self.frags.shape
        self.h = C.c_void_p(lib().wz_create_q(
            _p(self.affs), C.c_int(dt), _p(self.frags), C.c_int64(Z), C.c_int64(Y), C.c_int64(X),
            C.c_int(queue_bins), C.c_int({"faithful": 0, "canonical": 1}[stats_mode]),
            C.c_int(1 if keep_cheaper else 0), C.c_int(int(quantile)), C.c_int(1 if initmax else 0)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().wz_free(self.h)
            self.h = None

    @property
    def num_edges(self):
        return int(lib().wz_num_edges(self.h))

    def merge_until(self, threshold):
        n = int(lib().wz_merge_until(self.h, C.c_float(threshold)))
        a = np.empty(n, np.uint64)
        b = np.empty(n, np.uint64)
        c = np.empty(n, np.uint64)
        s = np.empty(n, np.float32)
        if n:
            lib().wz_history(self.h, _p(a), _p(b), _p(c), _p(s))
        return a, b, c, s

    def region_graph(self):
        E = self.num_edges
        u = np.empty(E, np.uint64)
        v = np.empty(E, np.uint64)
        s = np.empty(E, np.float32)
        sm = np.empty(E, np.float32)
        cn = np.empty(E, np.uint64)
        k = int(lib().wz_region_graph(self.h, _p(u), _p(v), _p(s), _p(sm), _p(cn)))
        return u[:k], v[:k], s[:k], sm[:k], cn[:k]

    def edge_stats(self):
        E = self.num_edges
        u = np.empty(E, np.uint64)
        v = np.empty(E, np.uint64)
        isum = np.empty(E, np.int64)
        cnt = np.empty(E, np.uint64)
        fsum = np.empty(E, np.float32)
        lib().wz_edge_stats(self.h, _p(u), _p(v), _p(isum), _p(cnt), _p(fsum))
        return u, v, isum, cnt, fsum

    def segmentation(self):
        seg = np.empty(self.frags.shape, np.uint64)
        lib().wz_segmentation(self.h, _p(self.frags), C.c_int64(self.frags.size), _p(seg))
        return seg

    def counters(self):
        out = np.zeros(3, np.uint64)
        lib().wz_counters(self.h, _p(out))
        return dict(pops=int(out[0]), stale=int(out[1]), deleted=int(out[2]))


def connected_components(nodes, edges, scores, threshold):
    """funlib.segment.graphs.impl.connected_components restated (post/watershed.py:182)."""
    nodes = np.ascontiguousarray(nodes, dtype=np.uint64)
    edges = np.ascontiguousarray(edges, dtype=np.uint64).reshape(-1, 2)
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    comp = np.empty(nodes.shape, np.uint64)
    lib().fs_connected_components(_p(nodes), C.c_int64(nodes.size), _p(edges), C.c_int64(edges.shape[0]),
                                  _p(scores), C.c_float(threshold), _p(comp))
    return comp
