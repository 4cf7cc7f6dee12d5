"""ORACLE (test infrastructure): the mutex-watershed path of `bs segment --mws`, restated.

* mwatershed_from_affinities: post/mws.py:12-59 line by line (shift = noise + (gauss - affs) + bias; mws.agglom on
  float64).  The reference's noise is an UNSEEDED np.random.randn: parity is impossible there, so `noise_seed` selects a
  seeded counter-based stand-in (sum of four 16-bit uniforms from a SplitMix64 hash of (seed, channel, voxel), unit
  variance) that the CUDA path (csrc/mws.cu) implements bit for bit.  randomized_strides draws an unseeded random subset of
  the stride lattice upstream: not restated, must be False.
* simple_mutex: post/watershed_mutex.py:177-291 on in-memory arrays (normalise, mask, fragments, remove_debris).
* mwatershed.agglom itself: oracle/csrc/mws_restated.cpp (PARITY UNPINNED, declared tie rule D4).
* volara_pipeline: post/watershed_mutex.py:8-174 on in-memory arrays.  The four tasks it runs are volara's own classes
  (volara < 1.0.3, third-party, not in the reference tree): restated from memory of the upstream source, PARITY UNPINNED.
    ExtractFrags  = the skeleton the reference's own WatershedFrags was copied from (watershed_frags.py:196-246: read ROI,
                    normalise, mask, fragments, filter_fragments, remove_debris, crop, skimage label, id bump, nodes) with
                    mwatershed.agglom(affs + shift, offsets=neighborhood, strides) on ALL channels as the fragmenter;
    AffAgglom     = per block, read ROI: for every offset c and voxel p with p + offset_c inside the read ROI, a pair of
                    different non-zero fragments adds affs[c][p] to edge (min, max); attribute "zyx_aff" = mean over all
                    contributions of all offsets; an edge is persisted by the block whose write ROI holds node min(u, v) (U9);
    GraphMWS      = global: w = weight * zyx_aff + bias for every edge, edges sorted by |w| descending (equal |w|: ascending
                    (u, v) -- declared, upstream it is the database's row order), mwatershed.cluster((w > 0, u, v)) -> LUT;
    Relabel       = replace_values(frags, lut).
"""
import numpy as np
from scipy.ndimage import gaussian_filter

from . import native as on
from bootstrapper_b200.synth import _hash

NOISE_K = np.sqrt(3.0) / 65536.0


def seeded_noise(shape, seed):
    """float64 array (C, Z, Y, X): the seeded stand-in for np.random.randn(*shape) (see module docstring)"""
    C = shape[0]
    V = int(np.prod(shape[1:]))
    p = np.arange(V, dtype=np.int64)
    out = np.empty((C, V), dtype=np.float64)
    for c in range(C):
        h = _hash(seed, np.full(V, c, dtype=np.int64), p, np.full(V, 11, dtype=np.int64))
        s = ((h & np.uint64(0xFFFF)).astype(np.int64) + ((h >> np.uint64(16)) & np.uint64(0xFFFF)).astype(np.int64)
             + ((h >> np.uint64(32)) & np.uint64(0xFFFF)).astype(np.int64) + (h >> np.uint64(48)).astype(np.int64))
        out[c] = (s - 131070).astype(np.float64) * NOISE_K
    return out.reshape(shape)


def mwatershed_from_affinities(affs, neighborhood, bias, sigma=None, noise_eps=None, strides=None, randomized_strides=False,
                               noise_seed=0, counters=None):
    """post/mws.py:12-59"""
    assert not randomized_strides, "randomized_strides is an unseeded RNG upstream"
    if sigma is not None:
        sigma = (0, *sigma)
    shift = np.zeros_like(affs)
    if noise_eps is not None:
        shift += seeded_noise(affs.shape, noise_seed) * noise_eps
    if sigma is not None:
        shift += gaussian_filter(affs, sigma=sigma) - affs
    shift += np.array([bias]).reshape((-1, *((1,) * (len(affs.shape) - 1))))
    return on.mws_agglom((affs + shift).astype(np.float64), neighborhood, strides, counters=counters)


def remove_small_objects(x, min_size):
    """skimage.morphology.remove_small_objects on a label array (U3): labels with fewer than min_size voxels -> 0"""
    out = x.copy()
    ids, counts = np.unique(x, return_counts=True)
    small = ids[(counts < min_size) & (ids != 0)]
    out[np.isin(x, small)] = 0
    return out


def simple_mutex(affs, params, mask=None, noise_seed=0):
    """post/watershed_mutex.py:177-291 on in-memory arrays: affs (C, Z, Y, X) uint8 or float.
    Returns dict(fragments, seg)."""
    neighborhood, bias = params["aff_neighborhood"], params["bias"]
    assert len(neighborhood) == affs.shape[0] == len(bias)
    if affs.dtype == np.uint8:
        affs_data = affs.astype(np.float64) / 255.0
    else:
        affs_data = affs.astype(np.float64)
    if mask is not None:
        affs_data *= (mask > 0).astype(np.uint8)
    frags = mwatershed_from_affinities(affs_data, neighborhood, bias, params.get("sigma"), params.get("noise_eps"),
                                       params.get("strides"), params.get("randomized_strides", False), noise_seed=noise_seed)
    seg = frags
    rd = params.get("remove_debris", 0)
    if rd and rd > 0:
        seg = remove_small_objects(frags.astype(np.int64), rd).astype(frags.dtype)
    return dict(fragments=frags, seg=seg)


# --------------------------------------------------------------------------- blockwise (volara tasks restated)
def mws_cluster(n_nodes, edges):
    """mwatershed.cluster on a graph: edges [(attractive, u, v)] in visiting order, node indices 0..n-1.
    Returns the root (smallest index) of every node's cluster."""
    parent = list(range(n_nodes))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    mutex = [set() for _ in range(n_nodes)]
    for attractive, u, v in edges:
        a, b = find(u), find(v)
        if a == b:
            continue
        if attractive:
            if b in mutex[a]:
                continue
            if len(mutex[a]) < len(mutex[b]):          # move the smaller mutex set
                a, b = b, a
            for m in mutex[b]:
                mutex[m].discard(b)
                mutex[m].add(a)
                mutex[a].add(m)
            mutex[b] = set()
            parent[b] = a
        else:
            mutex[a].add(b)
            mutex[b].add(a)
    roots = np.array([find(i) for i in range(n_nodes)], dtype=np.int64)
    # label every cluster by its smallest member
    smallest = np.full(n_nodes, n_nodes, dtype=np.int64)
    np.minimum.at(smallest, roots, np.arange(n_nodes))
    return smallest[roots]


def aff_agglom_in_block(block, affs, frags, rag, roi_offset, neighborhood, zyx):
    """volara AffAgglom.process_block restated (module docstring).  frags: task-ROI-sized array at roi_offset; zyx collects
    {(u, v): [integer sum, count]} of the block before the ownership rule, rag.edges receives the means of the owned edges."""
    import oracle.blockwise as ob
    affs_data = ob.to_ndarray(affs, block.read_offset, block.read_shape, 0)
    fo = [r - o for r, o in zip(block.read_offset, roi_offset)]
    fr = ob.to_ndarray(frags, fo, block.read_shape, 0)
    shape = fr.shape
    acc = {}
    for c, off in enumerate(neighborhood):
        base = tuple(slice(max(0, -o), min(s, s - o)) for o, s in zip(off, shape))
        shifted = tuple(slice(max(0, o), min(s, s + o)) for o, s in zip(off, shape))
        if any(sl.stop <= sl.start for sl in base):
            continue
        f1, f2, a = fr[base], fr[shifted], affs_data[c][base]
        m = (f1 != f2) & (f1 > 0) & (f2 > 0)
        lo, hi = np.minimum(f1[m], f2[m]), np.maximum(f1[m], f2[m])
        if affs_data.dtype == np.uint8:
            val = a[m].astype(np.int64)
        else:
            val = np.rint(np.ldexp(a[m].astype(np.float64), 38)).astype(np.int64)
        if lo.size == 0:
            continue
        pairs, inv = np.unique(np.stack([lo, hi], 1), axis=0, return_inverse=True)
        sums = np.zeros(len(pairs), dtype=np.int64)
        np.add.at(sums, inv.ravel(), val)
        cnts = np.bincount(inv.ravel(), minlength=len(pairs))
        for (u, v), sm, cn in zip(pairs, sums, cnts):
            e = acc.setdefault((int(u), int(v)), [0, 0])
            e[0] += int(sm)
            e[1] += int(cn)
    wlo = np.array(block.write_offset)
    whi = wlo + np.array(block.write_shape)
    for (u, v), (sm, cn) in acc.items():
        pos = rag.node_pos.get(u)
        if pos is None or not (np.all(np.array(pos) >= wlo) and np.all(np.array(pos) < whi)):
            continue
        if affs_data.dtype == np.uint8:
            mean = np.float32(np.float64(sm) / 255.0 / np.float64(cn))
        else:
            mean = np.float32(np.ldexp(np.float64(sm), -38) / np.float64(cn))
        rag.edges[(u, v)] = float(mean)
    if zyx is not None:
        zyx[block.block_id] = acc


def volara_pipeline(affs, params, block_size, context=None, mask=None, noise_seed=0, roi=None):
    """post/watershed_mutex.py:8-174 on in-memory arrays: affs (C, Z, Y, X) uint8 / float32.
    Returns dict(fragments, rag, lut (2, N), seg)."""
    import oracle.blockwise as ob
    from bootstrapper_b200.synth import block_seed
    neighborhood, bias = params["aff_neighborhood"], params["bias"]
    assert len(neighborhood) == affs.shape[0] == len(bias)
    assert params.get("sigma") is None and not params.get("randomized_strides", False)
    weight, gbias = tuple(params.get("global_bias", [1.0, -0.5]))
    filt = params.get("filter_fragments") or 0.0
    debris = params.get("remove_debris", 0) or 0
    vol = affs.shape[1:]
    roi_offset, roi_shape = roi if roi is not None else ((0, 0, 0), vol)
    if context is None:
        context = tuple(max(1, s // 8) for s in block_size)
    blocks = ob.enumerate_blocks(roi_offset, roi_shape, block_size, context)
    frags = np.zeros(roi_shape, dtype=np.uint64)
    rag = ob.Rag()

    def fragmenter(affs_data, block):
        f = mwatershed_from_affinities(affs_data, neighborhood, bias, None, params.get("noise_eps"), params.get("strides"), False,
                                       noise_seed=block_seed(noise_seed, block.block_id))
        if filt > 0:
            ob.filter_avg_fragments(affs_data, f, filt)
        if debris > 0:
            f = remove_small_objects(f.astype(np.int64), debris).astype(f.dtype)
        return f

    for blk in blocks:
        ob.watershed_in_block(blk, affs, frags, rag, {}, roi_offset, block_size, mask, fragmenter=fragmenter)
    for blk in blocks:
        aff_agglom_in_block(blk, affs, frags, rag, roi_offset, neighborhood, None)
    nodes = np.array(sorted(rag.node_pos), dtype=np.uint64)
    index = {int(n): i for i, n in enumerate(nodes)}
    items = sorted(rag.edges.items())
    scored = [(weight * np.float64(np.float32(s)) + gbias, u, v) for (u, v), s in items]
    order = sorted(range(len(scored)), key=lambda i: -abs(scored[i][0]))         # stable: equal |w| keep (u, v) order
    edges = [(scored[i][0] > 0, index[scored[i][1]], index[scored[i][2]]) for i in order]
    roots = mws_cluster(len(nodes), edges)
    lut = np.stack([nodes, nodes[roots]]) if len(nodes) else np.zeros((2, 0), np.uint64)
    seg = np.zeros_like(frags)
    if len(nodes):
        idx = np.searchsorted(nodes, frags)
        idx[idx >= len(nodes)] = 0
        hit = (frags > 0) & (nodes[idx] == frags)
        seg[hit] = lut[1][idx[hit]]
    return dict(fragments=frags, rag=rag, lut=lut, seg=seg, blocks=blocks)
