"""ctypes binding of libbsnative.so (the C ABI declared in include/bsnative.h).

The library is built in-tree by ``bootstrapper_b200/csrc/build.py`` (``__graft_entry__.build()``).
There is no CPU fallback: if the shared library is missing, or a call is made without a CUDA
device, this module raises.
"""
import ctypes as C
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbsnative.so")

BS_DTYPE_U8, BS_DTYPE_F32 = 0, 1
BS_ERR_OVERFLOW = -3
SIGMA_MAXW = 129


def gaussian_weights(sigma, truncate=4.0):
    """The 1-D kernel scipy.ndimage.gaussian_filter builds for one axis (order 0): radius = int(truncate * sigma + 0.5),
    weights exp(-x^2 / (2 sigma^2)) normalised to sum 1 (float64).  Returns (radius, weights); radius -1 when scipy skips
    the axis (sigma <= 1e-15)."""
    sd = float(sigma)
    if not sd > 1e-15:
        return -1, np.zeros(0)
    radius = int(truncate * sd + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sd * sd) * x ** 2)
    return radius, phi / phi.sum()


class BsError(RuntimeError):
    pass


class WsConfig(C.Structure):
    """bs_ws_config (include/bsnative.h)."""
    _fields_ = [
        ("vol_shape", C.c_int32 * 3), ("roi_offset", C.c_int32 * 3), ("roi_shape", C.c_int32 * 3),
        ("block_size", C.c_int32 * 3), ("context", C.c_int32 * 3),
        ("aff_dtype", C.c_int32), ("n_channels", C.c_int32), ("fragments_in_xy", C.c_int32),
        ("min_seed_distance", C.c_int32), ("remove_debris", C.c_int32), ("queue_bins", C.c_int32),
        ("keep_cheaper", C.c_int32), ("crop_relabel", C.c_int32), ("block_begin", C.c_int32),
        ("block_end", C.c_int32), ("win_z0", C.c_int32), ("win_z", C.c_int32), ("filter_fragments", C.c_double), ("max_batch_voxels", C.c_int64),
        ("has_bias", C.c_int32), ("has_seed_eps", C.c_int32), ("bias", C.c_double * 3), ("seed_eps", C.c_double),
        ("has_sigma", C.c_int32), ("sigma_radius", C.c_int32 * 3), ("sigma_w", (C.c_double * SIGMA_MAXW) * 3),
        ("block_index_offset", C.c_int32 * 3), ("has_noise", C.c_int32), ("noise_eps", C.c_double), ("noise_seed", C.c_uint64),
    ]


EXPORTS = [
    "bs_last_error", "bs_launch_count", "bs_version", "bs_config_size", "bs_plan_create", "bs_plan_destroy", "bs_plan_num_blocks",
    "bs_plan_block_info", "bs_plan_set_owned", "bs_stage1_fragments", "bs_stage1_num_nodes", "bs_stage1_get_nodes",
    "bs_stage1_block_counts", "bs_stage1_set_block_counts", "bs_plan_node_ids", "bs_stage2_agglomerate", "bs_stage2_agglomerate_until", "bs_stage1_from_labels", "bs_stage2_num_edges",
    "bs_stage2_get_edges", "bs_waterz_segment", "bs_waterz_segment_quantile", "bs_cc_affs", "bs_mws_agglom", "bs_mws_agglom_blocks", "bs_aff_agglom", "bs_graph_mws", "bs_aff_errors", "bs_label_stats", "bs_shift_affinities", "bs_connected_components", "bs_stage3_components", "bs_relabel", "bs_stage3_relabel", "bs_stage3_dense_fragments", "bs_expand_compact", "bs_watershed_from_affinities",
    "bs_synth_affs", "bs_debug_fetch", "bs_set_debug", "bs_set_profiling", "bs_get_profile",
    "bs_release_scratch", "bs_set_flood_version", "bs_set_front_version", "bs_set_agglom_version", "bs_dbg_scan_u32", "bs_dbg_scan_u8", "bs_dbg_sort_pairs",
]

_lib = None


def lib():
    """Load libbsnative.so; fail loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BsError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback for the bs segment hot path)")
        _lib = C.CDLL(LIB_PATH)
        _lib.bs_last_error.restype = C.c_char_p
        _lib.bs_launch_count.restype = C.c_ulonglong
        _lib.bs_config_size.restype = C.c_ulonglong
        for name in EXPORTS:
            getattr(_lib, name)
        if _lib.bs_config_size() != C.sizeof(WsConfig):
            raise BsError(f"bs_ws_config layout mismatch: library {_lib.bs_config_size()} bytes, binding {C.sizeof(WsConfig)}")
        # kernel-variant switches for experiments (see include/bsnative.h)
        if os.environ.get("BS_FLOOD_VERSION"):
            _lib.bs_set_flood_version(C.c_int(int(os.environ["BS_FLOOD_VERSION"])))
        if os.environ.get("BS_FRONT_VERSION"):
            _lib.bs_set_front_version(C.c_int(int(os.environ["BS_FRONT_VERSION"])))
        if os.environ.get("BS_AGGLOM_VERSION"):
            _lib.bs_set_agglom_version(C.c_int(int(os.environ["BS_AGGLOM_VERSION"])))
    return _lib


def _check(rc):
    if rc != 0:
        raise BsError(f"libbsnative error {rc}: {lib().bs_last_error().decode()}")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev(t, dtype=None):
    if not t.is_cuda:
        raise BsError("libbsnative needs CUDA tensors (no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise BsError(f"expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise BsError("tensor must be contiguous")
    return C.c_void_p(t.data_ptr())


def _aff_dtype(t):
    if t.dtype == torch.uint8:
        return BS_DTYPE_U8
    if t.dtype == torch.float32:
        return BS_DTYPE_F32
    raise BsError(f"affinities must be uint8 or float32, got {t.dtype}")


def launch_count():
    return int(lib().bs_launch_count())


def set_debug(on):
    _check(lib().bs_set_debug(C.c_int(1 if on else 0)))


def set_flood_version(v):
    _check(lib().bs_set_flood_version(C.c_int(int(v))))


def set_front_version(v):
    _check(lib().bs_set_front_version(C.c_int(int(v))))


def set_agglom_version(v):
    _check(lib().bs_set_agglom_version(C.c_int(int(v))))


def release_scratch():
    _check(lib().bs_release_scratch())


def set_profiling(on):
    _check(lib().bs_set_profiling(C.c_int(1 if on else 0)))


def get_profile():
    names = C.create_string_buffer(4096)
    ms = (C.c_float * 64)()
    n = C.c_int(0)
    _check(lib().bs_get_profile(names, 4096, ms, 64, C.byref(n)))
    keys = names.value.decode().split(";")[: n.value]
    return {k: float(ms[i]) for i, k in enumerate(keys)}


class Plan:
    """Geometry + results of one blockwise `bs segment --ws` run on this rank (bs_plan)."""

    def __init__(self, vol_shape, block_size, context, aff_dtype, roi_offset=None, roi_shape=None, n_channels=3,
                 fragments_in_xy=True, min_seed_distance=10, filter_fragments=0.1, remove_debris=64,
                 queue_bins=256, keep_cheaper=True, block_begin=-1, block_end=-1, max_batch_voxels=0, win_z0=0, win_z=0,
                 bias=None, seed_eps=None, sigma=None, block_index_offset=None, noise_eps=None, noise_seed=0):
        """block_index_offset: absolute offset of the task ROI in voxels (dataset offset / voxel_size + roi_offset), which
        daisy's block ids are numbered from; default: roi_offset, i.e. a dataset at world offset 0."""
        cfg = WsConfig()
        roi_offset = roi_offset if roi_offset is not None else (0, 0, 0)
        roi_shape = roi_shape if roi_shape is not None else vol_shape
        for d in range(3):
            cfg.vol_shape[d] = int(vol_shape[d])
            cfg.roi_offset[d] = int(roi_offset[d])
            cfg.roi_shape[d] = int(roi_shape[d])
            cfg.block_size[d] = int(block_size[d])
            cfg.context[d] = int(context[d])
            cfg.block_index_offset[d] = int((block_index_offset if block_index_offset is not None else roi_offset)[d])
        cfg.aff_dtype = aff_dtype
        cfg.n_channels = n_channels
        cfg.fragments_in_xy = 1 if fragments_in_xy else 0
        cfg.min_seed_distance = int(min_seed_distance)
        cfg.remove_debris = int(remove_debris or 0)
        cfg.queue_bins = int(queue_bins)
        cfg.keep_cheaper = 1 if keep_cheaper else 0
        cfg.crop_relabel = 1
        cfg.block_begin = int(block_begin)
        cfg.block_end = int(block_end)
        cfg.win_z0 = int(win_z0)
        cfg.win_z = int(win_z)
        cfg.filter_fragments = float(filter_fragments or 0.0)
        cfg.max_batch_voxels = int(max_batch_voxels)
        cfg.has_bias = 0 if bias is None else 1
        if bias is not None:   # watershed_frags.py:123-129: a scalar bias applies to every channel
            b = list(bias) if isinstance(bias, (list, tuple)) else [bias] * 3
            if len(b) != 3:
                raise BsError("bias must be a scalar or have one entry per affinity channel (3)")
            for d in range(3):
                cfg.bias[d] = float(b[d])
        cfg.has_seed_eps = 0 if seed_eps is None else 1
        cfg.seed_eps = float(seed_eps or 0.0)
        cfg.has_noise = 0 if noise_eps is None else 1     # watershed_frags.py:119-120, seeded (see bsnative.h)
        cfg.noise_eps = float(noise_eps or 0.0)
        cfg.noise_seed = int(noise_seed)
        cfg.has_sigma = 0
        if sigma is not None:   # watershed_frags.py:121-122: gaussian_filter(affs, sigma=(0, *sigma))
            if len(sigma) != 3:
                raise BsError("sigma must have one entry per spatial axis (z, y, x)")
            for d in range(3):
                r, w = gaussian_weights(sigma[d])
                if 2 * r + 1 > SIGMA_MAXW:
                    raise BsError(f"sigma {sigma[d]} is too large (kernel radius {r} > {(SIGMA_MAXW - 1) // 2})")
                cfg.sigma_radius[d] = r
                for i, v in enumerate(w):
                    cfg.sigma_w[d][i] = float(v)
                if r >= 0:
                    cfg.has_sigma = 1
        self.cfg = cfg
        self.roi_shape = tuple(int(v) for v in roi_shape)
        if win_z > 0:
            self.roi_shape = (int(win_z),) + self.roi_shape[1:]
        self._h = C.c_void_p()
        _check(lib().bs_plan_create(C.byref(cfg), C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value and _lib is not None:   # module globals may be gone at exit
            _lib.bs_plan_destroy(self._h)
            self._h = None

    # ---- geometry
    def num_blocks(self):
        a, b = C.c_int64(), C.c_int64()
        _check(lib().bs_plan_num_blocks(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def block_info(self):
        n, _ = self.num_blocks()
        ids = np.zeros(n, np.int64)
        wo = np.zeros((n, 3), np.int32)
        ws = np.zeros((n, 3), np.int32)
        _check(lib().bs_plan_block_info(self._h, ids.ctypes.data_as(C.c_void_p), wo.ctypes.data_as(C.c_void_p),
                                        ws.ctypes.data_as(C.c_void_p)))
        return ids, wo, ws

    def set_owned(self, indices):
        idx = np.ascontiguousarray(indices, dtype=np.int32)
        _check(lib().bs_plan_set_owned(self._h, idx.ctypes.data_as(C.c_void_p), C.c_int64(idx.size)))

    # ---- stage 1
    def fragments(self, affs, frags_out=None, mask=None):
        """WatershedFrags over all owned blocks.  Returns the uint64 fragments tensor (roi_shape),
        stored as torch.int64 bit patterns."""
        if frags_out is None:
            frags_out = torch.zeros(self.roi_shape, dtype=torch.int64, device=affs.device)
        vol = tuple(int(self.cfg.vol_shape[d]) for d in range(3))
        if self.cfg.win_z > 0:
            vol = (int(self.cfg.win_z),) + vol[1:]
        if tuple(affs.shape[1:]) != vol:
            raise BsError(f"affinities have spatial shape {tuple(affs.shape[1:])}, the plan was made for {vol}")
        if mask is not None and tuple(mask.shape) != vol:
            raise BsError(f"mask shape {tuple(mask.shape)} must equal the affinities' spatial shape {vol} "
                          "(crop / zero-pad it onto the affinity grid first)")
        _check(lib().bs_stage1_fragments(self._h, _dev(affs), _dev(mask, torch.uint8) if mask is not None else None,
                                         _dev(frags_out, torch.int64), _stream()))
        return frags_out

    def num_nodes(self):
        n = C.c_int64()
        _check(lib().bs_stage1_num_nodes(self._h, C.byref(n)))
        return n.value

    def nodes(self, device):
        n = self.num_nodes()
        ids = torch.empty(n, dtype=torch.int64, device=device)
        pos = torch.empty((n, 3), dtype=torch.int32, device=device)
        sizes = torch.empty(n, dtype=torch.int32, device=device)
        _check(lib().bs_stage1_get_nodes(self._h, _dev(ids), _dev(pos), _dev(sizes), _stream()))
        return ids, pos, sizes

    def block_counts(self):
        n, _ = self.num_blocks()
        c = np.zeros(n, np.int64)
        _check(lib().bs_stage1_block_counts(self._h, c.ctypes.data_as(C.c_void_p)))
        return c

    def set_block_counts(self, counts):
        c = np.ascontiguousarray(counts, dtype=np.int64)
        _check(lib().bs_stage1_set_block_counts(self._h, c.ctypes.data_as(C.c_void_p)))

    def node_ids(self, device):
        """ascending ids of all fragments of the task (needs every block's count: single rank, or after
        set_block_counts)"""
        n = C.c_int64()
        _check(lib().bs_plan_node_ids(self._h, None, C.byref(n), _stream()))
        ids = torch.empty(n.value, dtype=torch.int64, device=device)
        _check(lib().bs_plan_node_ids(self._h, _dev(ids), C.byref(n), _stream()))
        return ids

    # ---- stage 2
    def agglomerate(self, affs, frags):
        _check(lib().bs_stage2_agglomerate(self._h, _dev(affs), _dev(frags, torch.int64), _stream()))

    def aff_agglom(self, affs, frags, offsets):
        """volara AffAgglom(scores={"zyx_aff": neighborhood}) for the owned blocks (bs_aff_agglom): mean affinity over all
        offsets between every pair of fragments that touch through one; results through edges()"""
        off = np.ascontiguousarray(np.asarray(offsets, dtype=np.int32).reshape(-1, 3))
        if off.shape[0] != affs.shape[0]:
            raise BsError("Number of offsets must match number of affinities channels")
        _check(lib().bs_aff_agglom(self._h, _dev(affs), _dev(frags, torch.int64), C.c_int(int(affs.shape[0])), off.ctypes.data_as(C.c_void_p),
                                   _stream()))

    def agglomerate_until(self, affs, frags, threshold):
        """waterz mergeUntil(threshold) instead of the blockwise 1.0 (epsilon_agglomerate, watershed_frags.py:158-176): the
        edges of merged pairs carry their merge score, all others NaN"""
        _check(lib().bs_stage2_agglomerate_until(self._h, _dev(affs), _dev(frags, torch.int64), C.c_float(float(threshold)), _stream()))

    def fragments_from_labels(self, affs, labels, n_labels, frags_out, mask=None):
        """the back half of WatershedFrags on given fragments (bs_stage1_from_labels): labels int32, one read-ROI volume per
        owned block packed back to back, values 0 or 1..n_labels"""
        _check(lib().bs_stage1_from_labels(self._h, _dev(affs), _dev(mask, torch.uint8) if mask is not None else None,
                                           _dev(labels, torch.int32), C.c_int64(int(n_labels)), _dev(frags_out, torch.int64), _stream()))
        return frags_out

    def num_edges(self):
        n = C.c_int64()
        _check(lib().bs_stage2_num_edges(self._h, C.byref(n)))
        return n.value

    def edges(self, device):
        n = self.num_edges()
        u = torch.empty(n, dtype=torch.int64, device=device)
        v = torch.empty(n, dtype=torch.int64, device=device)
        s = torch.empty(n, dtype=torch.float32, device=device)
        _check(lib().bs_stage2_get_edges(self._h, _dev(u), _dev(v), _dev(s), _stream()))
        return u, v, s

    # ---- single-shot path (waterz with the non-discretised queue)
    def waterz_segment(self, affs, frags, thresholds, outs=None, quantile=0, init_with_max=False):
        """waterz.agglomerate(affs, thresholds, fragments, scoring_function) on a single-block plan (post/watershed.py:333-340);
        quantile = 0: OneMinus<MeanAffinity>, Q: OneMinus<HistogramQuantileAffinity<Q, 256, init_with_max>> (:232-244).
        Returns ([segmentation per ascending threshold], sorted thresholds, counters dict)."""
        thr = np.ascontiguousarray(sorted(float(t) for t in thresholds), dtype=np.float32)
        T = len(thr)
        if outs is None:
            outs = [torch.empty_like(frags) for _ in range(T)]
        for t in outs:
            _dev(t, torch.int64)
        sp = (C.c_void_p * T)(*[o.data_ptr() for o in outs])
        cnt = (C.c_uint32 * 4)()
        _check(lib().bs_waterz_segment_quantile(self._h, _dev(affs), _dev(frags, torch.int64), thr.ctypes.data_as(C.c_void_p),
                                                C.c_int(T), C.c_int(int(quantile)), C.c_int(1 if init_with_max else 0), sp, cnt, _stream()))
        return outs, [float(t) for t in thr], dict(pops=cnt[0], stale=cnt[1], deleted=cnt[2], merges=cnt[3])

    # ---- stage 3
    def components(self, nodes, edges_u, edges_v, scores, thresholds):
        """connected_components (post/watershed.py:182) for every threshold of the run in one pass over the edges.
        Returns {threshold: components tensor} (component id = smallest node id of the component)."""
        order = sorted(set(float(t) for t in thresholds))
        out = {}
        for i in range(0, len(order), 8):
            thr = np.ascontiguousarray(order[i:i + 8], dtype=np.float32)
            comps = [torch.empty_like(nodes) for _ in thr]
            cp = (C.c_void_p * len(thr))(*[c.data_ptr() for c in comps])
            m = edges_u.numel()
            _check(lib().bs_stage3_components(self._h, _dev(nodes, torch.int64), C.c_int64(nodes.numel()),
                                              _dev(edges_u, torch.int64) if m else None, _dev(edges_v, torch.int64) if m else None,
                                              _dev(scores, torch.float32) if m else None, C.c_int64(m),
                                              thr.ctypes.data_as(C.c_void_p), C.c_int(len(thr)), cp, _stream()))
            for t, c in zip(order[i:i + 8], comps):
                out[t] = c
        return {t: out[float(t)] for t in thresholds}

    def relabel(self, frags, comps, outs=None):
        """all thresholds in one pass; comps: list of (N,) LUT value tensors in ascending node-id order"""
        T = len(comps)
        if outs is None:
            outs = [torch.empty_like(frags) for _ in range(T)]
        cp = (C.c_void_p * T)(*[c.data_ptr() for c in comps])
        sp = (C.c_void_p * T)(*[o.data_ptr() for o in outs])
        for t in list(comps) + list(outs):
            _dev(t, torch.int64)
        _check(lib().bs_stage3_relabel(self._h, _dev(frags, torch.int64), C.c_int64(frags.numel()), cp, C.c_int(T), sp, _stream()))
        return outs

    def dense_fragments(self, frags, out=None):
        """compact form of a fragment volume: int32 plane of 1 + node rank (0 = background), see include/bsnative.h"""
        if out is None:
            out = torch.empty(frags.shape, dtype=torch.int32, device=frags.device)
        _check(lib().bs_stage3_dense_fragments(self._h, _dev(frags, torch.int64), C.c_int64(frags.numel()), _dev(out, torch.int32), _stream()))
        return out

    # ---- debug
    def debug_fetch(self, name, dtype):
        n = C.c_int64()
        _check(lib().bs_debug_fetch(self._h, name.encode(), None, C.byref(n)))
        out = np.empty(n.value, dtype=dtype)
        _check(lib().bs_debug_fetch(self._h, name.encode(), out.ctypes.data_as(C.c_void_p), C.byref(n)))
        return out


def connected_components(nodes, edges_u, edges_v, scores, threshold):
    """funlib.segment.graphs.impl.connected_components on the device (post/watershed.py:182)."""
    comp = torch.empty_like(nodes)
    _check(lib().bs_connected_components(_dev(nodes, torch.int64), C.c_int64(nodes.numel()),
                                         _dev(edges_u, torch.int64) if edges_u.numel() else None,
                                         _dev(edges_v, torch.int64) if edges_v.numel() else None,
                                         _dev(scores, torch.float32) if scores.numel() else None,
                                         C.c_int64(edges_u.numel()), C.c_float(threshold), _dev(comp), _stream()))
    return comp


def relabel(frags, lut_keys, lut_vals, out=None):
    """volara Relabel / replace_values on the device (post/watershed.py:192-202)."""
    if out is None:
        out = torch.empty_like(frags)
    _check(lib().bs_relabel(_dev(frags, torch.int64), C.c_int64(frags.numel()),
                            _dev(lut_keys, torch.int64) if lut_keys.numel() else None,
                            _dev(lut_vals, torch.int64) if lut_vals.numel() else None,
                            C.c_int64(lut_keys.numel()), _dev(out, torch.int64), _stream()))
    return out


def expand_compact(dense, node_ids, luts, frags_out=None, segs_out=None, threads=None):
    """HOST decoder of the compact result form (bs_expand_compact): dense int32 plane + node-id table + LUT rows (all CPU
    tensors) -> the uint64 fragments / segmentations, written into the given CPU tensors (allocated when None)."""
    for t in [dense, node_ids] + list(luts):
        if t.is_cuda or not t.is_contiguous():
            raise BsError("expand_compact works on contiguous host tensors")
    T = len(luts)
    if frags_out is None:
        frags_out = torch.empty(dense.shape, dtype=torch.int64)
    if segs_out is None:
        segs_out = [torch.empty(dense.shape, dtype=torch.int64) for _ in range(T)]
    lp = (C.c_void_p * max(T, 1))(*[l.data_ptr() for l in luts])
    sp = (C.c_void_p * max(T, 1))(*[o.data_ptr() for o in segs_out])
    _check(lib().bs_expand_compact(C.c_void_p(dense.data_ptr()), C.c_int64(dense.numel()), C.c_void_p(node_ids.data_ptr()),
                                   C.c_int64(node_ids.numel()), lp, C.c_int(T), C.c_void_p(frags_out.data_ptr()), sp,
                                   C.c_int(int(threads or os.cpu_count() or 1))))
    return frags_out, segs_out


def cc_affs(affs, threshold, remove_debris=0, mask=None):
    """post/connected_components.py cc_affs on one device array (C >= 3, Z, Y, X) uint8 / float32.
    Returns (fragments, segmentation after remove_debris, number of components)."""
    _, Z, Y, X = affs.shape
    frags = torch.empty((Z, Y, X), dtype=torch.int64, device=affs.device)
    seg = torch.empty_like(frags)
    n = C.c_int64()
    _check(lib().bs_cc_affs(_dev(affs), C.c_int(_aff_dtype(affs)), _dev(mask, torch.uint8) if mask is not None else None,
                            Z, Y, X, C.c_float(float(threshold)), C.c_int(int(remove_debris or 0)), _dev(frags), _dev(seg),
                            C.byref(n), _stream()))
    return frags, seg, n.value


def mws_agglom(affs, offsets, bias, strides=None, mask=None, noise_eps=None, noise_seed=0, remove_debris=0):
    """mutex watershed of post/mws.py on one device array (C, Z, Y, X) uint8 / float32 (bs_mws_agglom).
    Returns (fragments, fragments after remove_debris, counters)."""
    Cn, Z, Y, X = affs.shape
    off = np.ascontiguousarray(np.asarray(offsets, dtype=np.int32).reshape(-1, 3))
    if off.shape[0] != Cn:
        raise BsError("Number of offsets must match number of affinities channels")
    b = np.ascontiguousarray(np.asarray(bias, dtype=np.float64).reshape(-1))
    if b.shape[0] != Cn:
        raise BsError("Number of biases must match number of affinities channels")
    st = None
    if strides is not None:
        st = np.ascontiguousarray(np.asarray(strides, dtype=np.int32).reshape(-1, 3))
        if st.shape[0] != Cn:
            raise BsError("Number of strides must match number of affinities channels")
    frags = torch.empty((Z, Y, X), dtype=torch.int64, device=affs.device)
    seg = torch.empty_like(frags)
    cnt = (C.c_int64 * 5)()
    _check(lib().bs_mws_agglom(_dev(affs), C.c_int(_aff_dtype(affs)), _dev(mask, torch.uint8) if mask is not None else None,
                               C.c_int(Cn), Z, Y, X, off.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p) if st is not None else None,
                               b.ctypes.data_as(C.c_void_p), C.c_double(float(noise_eps or 0.0)), C.c_ulonglong(int(noise_seed)),
                               C.c_int(int(remove_debris or 0)), _dev(frags), _dev(seg), cnt, _stream()))
    return frags, seg, dict(edges=cnt[0], merges=cnt[1], mutexes=cnt[2], blocked=cnt[3], rounds=cnt[4])


def _mws_arrays(Cn, offsets, bias, strides):
    off = np.ascontiguousarray(np.asarray(offsets, dtype=np.int32).reshape(-1, 3))
    if off.shape[0] != Cn:
        raise BsError("Number of offsets must match number of affinities channels")
    b = np.ascontiguousarray(np.asarray(bias, dtype=np.float64).reshape(-1))
    if b.shape[0] != Cn:
        raise BsError("Number of biases must match number of affinities channels")
    st = None
    if strides is not None:
        st = np.ascontiguousarray(np.asarray(strides, dtype=np.int32).reshape(-1, 3))
        if st.shape[0] != Cn:
            raise BsError("Number of strides must match number of affinities channels")
    return off, b, st


def mws_agglom_blocks(affs, n_blocks, offsets, bias, strides=None, noise_eps=None, block_seeds=None):
    """volara ExtractFrags' fragmenter for n_blocks read ROIs stacked along z in one device array (C, n_blocks * Z, Y, X)
    (bs_mws_agglom_blocks): one independent mutex watershed per block.  Returns (int32 labels (n_blocks * Z, Y, X): clusters
    numbered 1..n in the order of their first voxel in the stack, counters incl. n_labels)."""
    Cn, ZZ, Y, X = affs.shape
    if ZZ % n_blocks:
        raise BsError("stacked depth is not a multiple of the block count")
    off, b, st = _mws_arrays(Cn, offsets, bias, strides)
    seeds = None
    if noise_eps:
        seeds = np.ascontiguousarray(np.asarray(block_seeds, dtype=np.uint64).reshape(-1))
        if seeds.shape[0] != n_blocks:
            raise BsError("one noise seed per block is needed")
    labels = torch.empty((ZZ, Y, X), dtype=torch.int32, device=affs.device)
    cnt = (C.c_int64 * 6)()
    _check(lib().bs_mws_agglom_blocks(_dev(affs), C.c_int(_aff_dtype(affs)), None, C.c_int(Cn), C.c_int(int(n_blocks)), C.c_int(ZZ // n_blocks),
                                      C.c_int(Y), C.c_int(X), off.ctypes.data_as(C.c_void_p),
                                      st.ctypes.data_as(C.c_void_p) if st is not None else None, b.ctypes.data_as(C.c_void_p),
                                      C.c_double(float(noise_eps or 0.0)), seeds.ctypes.data_as(C.c_void_p) if seeds is not None else None,
                                      _dev(labels, torch.int32), cnt, _stream()))
    return labels, dict(edges=cnt[0], merges=cnt[1], mutexes=cnt[2], blocked=cnt[3], rounds=cnt[4], n_labels=cnt[5])


def graph_mws(nodes, edges_u, edges_v, scores, weight=1.0, bias=-0.5):
    """volara GraphMWS on the device (bs_graph_mws): mutex watershed over the fragment graph with w = weight * score + bias;
    edges must come sorted by (u, v) (the declared order of equal |w|).  Returns (clusters: smallest node id of every node's
    cluster, counters)."""
    out = torch.empty_like(nodes)
    cnt = (C.c_int64 * 5)()
    m = edges_u.numel()
    _check(lib().bs_graph_mws(_dev(nodes, torch.int64) if nodes.numel() else None, C.c_int64(nodes.numel()),
                              _dev(edges_u, torch.int64) if m else None, _dev(edges_v, torch.int64) if m else None,
                              _dev(scores, torch.float32) if m else None, C.c_int64(m), C.c_double(float(weight)), C.c_double(float(bias)),
                              _dev(out, torch.int64) if nodes.numel() else None, cnt, _stream()))
    return out, dict(edges=cnt[0], merges=cnt[1], mutexes=cnt[2], blocked=cnt[3], rounds=cnt[4])


def label_stats(seg, capacity=1 << 20):
    """ids (ascending), voxel counts, first / last z plane of every non-zero id of a CUDA label volume (Z,Y,X) int64.
    The table grows until it holds every id."""
    shape = (C.c_int32 * 3)(*[int(v) for v in seg.shape])
    while True:
        ids = torch.empty(capacity, dtype=torch.int64, device=seg.device)
        sizes = torch.empty(capacity, dtype=torch.int64, device=seg.device)
        zlo = torch.empty(capacity, dtype=torch.int32, device=seg.device)
        zhi = torch.empty(capacity, dtype=torch.int32, device=seg.device)
        n = C.c_int64()
        rc = lib().bs_label_stats(_dev(seg, torch.int64), shape, C.c_int64(capacity), _dev(ids), _dev(sizes), _dev(zlo), _dev(zhi),
                                  C.byref(n), _stream())
        if rc == BS_ERR_OVERFLOW and capacity < (1 << 29):
            capacity *= 4
            continue
        _check(rc)
        k = n.value
        return ids[:k], sizes[:k], zlo[:k], zhi[:k]


def aff_errors(seg, pred_affs, neighborhood, labels_mask=None, thresholds=(0.1, 1.0), return_seg_affs=True):
    """AddAffErrors.process on one device array (gp/add_aff_errors.py:128-183): seg (Z,Y,X) int64/uint64 ids, pred_affs
    (C,Z,Y,X) float32 or uint8, neighborhood C x 3 offsets.  Returns (seg_affs float32 or None, error_map float32,
    error_mask uint8)."""
    nh = np.ascontiguousarray(np.asarray(neighborhood, dtype=np.int32).reshape(-1, 3))
    Cn = nh.shape[0]
    if pred_affs.shape[0] != Cn or tuple(pred_affs.shape[1:]) != tuple(seg.shape):
        raise BsError("pred_affs must be (len(neighborhood),) + seg.shape")
    shape = (C.c_int32 * 3)(*[int(v) for v in seg.shape])
    seg_affs = torch.empty((Cn,) + tuple(seg.shape), dtype=torch.float32, device=seg.device) if return_seg_affs else None
    err = torch.empty(tuple(seg.shape), dtype=torch.float32, device=seg.device)
    emask = torch.empty(tuple(seg.shape), dtype=torch.uint8, device=seg.device)
    _check(lib().bs_aff_errors(_dev(seg, torch.int64), _dev(pred_affs), C.c_int(_aff_dtype(pred_affs)), C.c_int(Cn), shape,
                               nh.ctypes.data_as(C.c_void_p), _dev(labels_mask, torch.uint8) if labels_mask is not None else None,
                               C.c_float(float(thresholds[0])), C.c_float(float(thresholds[1])),
                               _dev(seg_affs) if seg_affs is not None else None, _dev(err), _dev(emask), _stream()))
    return seg_affs, err, emask


def shift_affinities(affs, mask=None, sigma=None, bias=None):
    """affs_data + shift of the single-shot paths (post/watershed.py:262-303, connected_components.py:52-77) as a
    float32 tensor (3, Z, Y, X); affs: CUDA tensor (C >= 3, Z, Y, X) uint8 or float32."""
    _, Z, Y, X = affs.shape
    out = torch.empty((3, Z, Y, X), dtype=torch.float32, device=affs.device)
    rad = (C.c_int32 * 3)(-1, -1, -1)
    wptr = (C.c_void_p * 3)()
    keep = []
    if sigma is not None:
        if len(sigma) != 3:
            raise BsError("sigma must have one entry per spatial axis (z, y, x)")
        for d in range(3):
            r, w = gaussian_weights(sigma[d])
            if 2 * r + 1 > SIGMA_MAXW:
                raise BsError(f"sigma {sigma[d]} is too large")
            rad[d] = r
            w = np.ascontiguousarray(w, dtype=np.float64)
            keep.append(w)
            wptr[d] = w.ctypes.data if r >= 0 else None
    b = None
    if bias is not None:
        bl = list(bias) if isinstance(bias, (list, tuple)) else [bias] * 3
        if len(bl) != 3:
            raise BsError("bias must be a scalar or have one entry per affinity channel (3)")
        b = (C.c_double * 3)(*[float(v) for v in bl])
    _check(lib().bs_shift_affinities(_dev(affs), C.c_int(_aff_dtype(affs)), _dev(mask, torch.uint8) if mask is not None else None,
                                     Z, Y, X, rad, wptr, b, _dev(out), _stream()))
    return out


def watershed_from_affinities(affs, fragments_in_xy, min_seed_distance):
    """post/ws.py:38 on one device array.  Returns (fragments int64 tensor, n)."""
    _, Z, Y, X = affs.shape
    out = torch.zeros((Z, Y, X), dtype=torch.int64, device=affs.device)
    n = C.c_int64()
    _check(lib().bs_watershed_from_affinities(_dev(affs), C.c_int(_aff_dtype(affs)), Z, Y, X,
                                              C.c_int(1 if fragments_in_xy else 0), C.c_int(int(min_seed_distance)),
                                              _dev(out), None, C.byref(n), _stream()))
    return out, n.value


def synth_affs(shape, seed=0, dtype=torch.uint8, offset=(0, 0, 0), device="cuda"):
    """Device generator, bit-identical to bootstrapper_b200.synth.synth_affs."""
    out = torch.empty((3,) + tuple(shape), dtype=dtype, device=device)
    sh = (C.c_int32 * 3)(*[int(v) for v in shape])
    of = (C.c_int32 * 3)(*[int(v) for v in offset])
    _check(lib().bs_synth_affs(_dev(out), C.c_int(_aff_dtype(out)), sh, of, None, C.c_uint64(seed), _stream()))
    return out
