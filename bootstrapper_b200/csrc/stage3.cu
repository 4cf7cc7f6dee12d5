// Stage 3: global thresholded connected components over the fragment graph, LUT, relabel.
//
// Replaces funlib.segment.graphs.impl.connected_components (post/watershed.py:177-182, U7: edges
// with score <= threshold, float32 compare), volara LUT + Relabel (post/watershed.py:187-202).
#include <algorithm>

#include "geom.h"

namespace bs {

static constexpr uint32_t NONE32 = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t find_sorted(const uint64_t *__restrict__ keys, uint32_t n, uint64_t id) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (keys[mid] < id)
            lo = mid + 1;
        else
            hi = mid;
    }
    return (lo < n && keys[lo] == id) ? lo : NONE32;
}

__global__ void k_iota(uint32_t *p, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

__global__ void k_cc_union(const uint64_t *__restrict__ nodes, uint32_t n, const uint64_t *__restrict__ eu,
                           const uint64_t *__restrict__ ev, const float *__restrict__ scores, size_t m, float thr,
                           uint32_t *parent) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m) return;
    if (scores && !(scores[e] <= thr)) return;   // NaN (merge_score NULL) never passes
    uint32_t a = find_sorted(nodes, n, eu[e]), b = find_sorted(nodes, n, ev[e]);
    if (a == NONE32 || b == NONE32 || a == b) return;
    uf_union(parent, a, b);
}

__global__ void k_cc_flatten(const uint64_t *__restrict__ nodes, uint32_t n, const uint32_t *parent, uint64_t *__restrict__ comp) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) comp[i] = nodes[uf_find(parent, i)];
}

int connected_components(const uint64_t *nodes, int64_t n, const uint64_t *eu, const uint64_t *ev, const float *scores,
                         int64_t m, float thr, uint64_t *comp, cudaStream_t s) {
    BS_ARG(n >= 0 && n < (1LL << 32) - 1 && m >= 0, "bs_connected_components: bad sizes");
    if (n == 0) return BS_OK;
    DevBuf parent;
    BS_TRY(parent.alloc(4 * (size_t)n, s));
    BS_LAUNCH(k_iota, cdiv(n, 256), 256, 0, s, parent.as<uint32_t>(), (uint32_t)n);
    if (m) BS_LAUNCH(k_cc_union, cdiv(m, 256), 256, 0, s, nodes, (uint32_t)n, eu, ev, scores, (size_t)m, thr, parent.as<uint32_t>());
    BS_LAUNCH(k_cc_flatten, cdiv(n, 256), 256, 0, s, nodes, (uint32_t)n, parent.as<uint32_t>(), comp);
    BS_CUDA(cudaGetLastError());
    return BS_OK;
}

// generic LUT relabel (volara Relabel / funlib replace_values): ids absent from the LUT are unchanged.
// Neighbouring voxels mostly carry the same id: the previous hit is reused before searching.
__global__ void __launch_bounds__(256) k_relabel(const uint64_t *__restrict__ frags, size_t n, const uint64_t *__restrict__ keys,
                                                 const uint64_t *__restrict__ vals, uint32_t k, uint64_t *__restrict__ seg) {
    uint64_t last_id = 0, last_val = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t id = frags[i];
        uint64_t out = id;
        if (id != 0) {
            if (id == last_id)
                out = last_val;
            else {
                uint32_t j = find_sorted(keys, k, id);
                if (j != NONE32) out = vals[j];
                last_id = id;
                last_val = out;
            }
        }
        seg[i] = out;
    }
}

int relabel(const uint64_t *frags, int64_t n, const uint64_t *keys, const uint64_t *vals, int64_t k, uint64_t *seg,
            cudaStream_t s) {
    BS_ARG(n >= 0 && k >= 0 && k < (1LL << 32) - 1, "bs_relabel: bad sizes");
    if (n == 0) return BS_OK;
    unsigned grid = (unsigned)std::min<size_t>(cdiv(n, 256), 148 * 16 * 8);
    BS_LAUNCH(k_relabel, grid, 256, 0, s, frags, (size_t)n, keys, vals, (uint32_t)k, seg);
    BS_CUDA(cudaGetLastError());
    return BS_OK;
}

// ---- Relabel of all thresholds in one pass: the LUT is indexed by the dense node number derived from the id
struct RelabelSet {
    const uint64_t *comp[8];
    uint64_t *seg[8];
    int T;
};

__global__ void __launch_bounds__(256) k_relabel_dense(const uint64_t *__restrict__ frags, size_t n, IdMap idm, uint32_t n_nodes,
                                                       RelabelSet rs) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t id = frags[i];
        uint32_t d = id_to_dense(idm, id);
        bool hit = d != NONE32 && d < n_nodes;
#pragma unroll
        for (int t = 0; t < 8; t++)
            if (t < rs.T) rs.seg[t][i] = hit ? rs.comp[t][d] : id;
    }
}

int relabel_dense(Plan &P, const uint64_t *frags, int64_t n, const uint64_t *const *comps, int T, uint64_t *const *segs,
                  cudaStream_t s) {
    BS_ARG(T >= 1 && T <= 8, "bs_stage3_relabel: 1..8 thresholds per call");
    if (n == 0) return BS_OK;
    const size_t nblocks = P.blocks.size();
    DevBuf d_c2d;
    IdMap idm;
    BS_TRY(plan_idmap(P, d_c2d, &idm, s));
    RelabelSet rs;
    rs.T = T;
    for (int t = 0; t < 8; t++) {
        rs.comp[t] = t < T ? comps[t] : nullptr;
        rs.seg[t] = t < T ? segs[t] : nullptr;
    }
    unsigned grid = (unsigned)std::min<size_t>(cdiv(n, 256), 148 * 16 * 8);
    BS_LAUNCH(k_relabel_dense, grid, 256, 0, s, frags, (size_t)n, idm, (uint32_t)P.block_nbase[nblocks], rs);
    BS_CUDA(cudaStreamSynchronize(s));   // c2d is a host-staged copy
    BS_CUDA(cudaGetLastError());
    return BS_OK;
}

// ---- thresholded connected components for all thresholds of a run in one pass over the edges: node numbers come
// from the id arithmetic of the plan (no search), one union-find forest per threshold
struct CcSet {
    float thr[8];
    uint64_t *comp[8];
    int T;
};

__global__ void k_cc_init_multi(uint32_t *__restrict__ parent, size_t n_total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_total) parent[i] = 0xFFFFFFFFu;   // "root" (own index implied)
}

__device__ __forceinline__ uint32_t ufm_find(const uint32_t *parent, uint32_t x) {
    uint32_t p = __ldcg(&parent[x]);
    while (p != NONE32) {
        x = p;
        p = __ldcg(&parent[x]);
    }
    return x;
}
__device__ __forceinline__ void ufm_union(uint32_t *parent, uint32_t a, uint32_t b) {
    for (;;) {
        a = ufm_find(parent, a);
        b = ufm_find(parent, b);
        if (a == b) return;
        if (a < b) {
            uint32_t t = a;
            a = b;
            b = t;
        }
        uint32_t old = atomicMin(&parent[a], b);   // a root holds NONE32: any index is smaller
        if (old == NONE32) return;
        a = old;
    }
}

__global__ void __launch_bounds__(256) k_cc_union_multi(const uint64_t *__restrict__ eu, const uint64_t *__restrict__ ev,
                                                        const float *__restrict__ scores, size_t m, IdMap idm, uint32_t n, CcSet cs,
                                                        uint32_t *__restrict__ parent) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m) return;
    const float sc = scores[e];
    if (!(sc <= cs.thr[cs.T - 1])) return;   // thresholds ascending; NaN (merge_score NULL) never passes
    const uint32_t a = id_to_dense(idm, eu[e]), b = id_to_dense(idm, ev[e]);
    if (a >= n || b >= n || a == b) return;
#pragma unroll
    for (int t = 0; t < 8; t++)
        if (t < cs.T && sc <= cs.thr[t]) ufm_union(parent + (size_t)t * n, a, b);
}

__global__ void __launch_bounds__(256) k_cc_flatten_multi(const uint64_t *__restrict__ nodes, uint32_t n, const uint32_t *__restrict__ parent,
                                                          CcSet cs) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int t = 0; t < 8; t++)
        if (t < cs.T) cs.comp[t][i] = nodes[ufm_find(parent + (size_t)t * n, i)];
}

// nodes: ascending ids of all fragments of the task (bs_plan_node_ids); thresholds ascending
int components_multi(Plan &P, const uint64_t *nodes, int64_t n, const uint64_t *eu, const uint64_t *ev, const float *scores,
                     int64_t m, const float *thresholds, int T, uint64_t *const *comps, cudaStream_t s) {
    BS_ARG(T >= 1 && T <= 8, "bs_stage3_components: 1..8 thresholds per call");
    BS_ARG(n == P.block_nbase[P.blocks.size()], "bs_stage3_components: node list does not match the plan's fragment counts");
    for (int t = 1; t < T; t++) BS_ARG(thresholds[t] >= thresholds[t - 1], "bs_stage3_components: thresholds must be ascending");
    if (n == 0) return BS_OK;
    const size_t nblocks = P.blocks.size();
    DevBuf d_c2d, parent;
    IdMap idm;
    BS_TRY(plan_idmap(P, d_c2d, &idm, s));
    CcSet cs;
    cs.T = T;
    for (int t = 0; t < 8; t++) {
        cs.thr[t] = t < T ? thresholds[t] : 0.f;
        cs.comp[t] = t < T ? comps[t] : nullptr;
    }
    BS_TRY(parent.alloc(4 * (size_t)n * T, s));
    BS_LAUNCH(k_cc_init_multi, cdiv((size_t)n * T, 256), 256, 0, s, parent.as<uint32_t>(), (size_t)n * T);
    if (m) BS_LAUNCH(k_cc_union_multi, cdiv((size_t)m, 256), 256, 0, s, eu, ev, scores, (size_t)m, idm, (uint32_t)n, cs, parent.as<uint32_t>());
    BS_LAUNCH(k_cc_flatten_multi, cdiv((size_t)n, 256), 256, 0, s, nodes, (uint32_t)n, parent.as<uint32_t>(), cs);
    BS_CUDA(cudaStreamSynchronize(s));   // c2d is a host-staged copy
    BS_CUDA(cudaGetLastError());
    return BS_OK;
}

// ---- node ids of the whole task from the per-block fragment counts (ids are 1..n per block + block_id * prod(block_size),
// watershed_frags.py:224; blocks ascending by id): the sorted key row of the fragment -> segment LUT
__global__ void k_node_ids(const long long *__restrict__ nbase, const long long *__restrict__ bid, int nblocks,
                           long long nvox_block, long long n, uint64_t *__restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo = 0, hi = nblocks - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (nbase[mid] <= i)
            lo = mid;
        else
            hi = mid - 1;
    }
    out[i] = (uint64_t)(i - nbase[lo] + 1) + (uint64_t)bid[lo] * (uint64_t)nvox_block;
}

// fragment ids -> dense node numbers + 1 (0: background / unknown id): the compact form of a fragment volume (4 bytes per
// voxel); together with the node-id table and one LUT row per threshold it determines fragments and all segmentations
__global__ void __launch_bounds__(256) k_dense_ids(const uint64_t *__restrict__ frags, size_t n, IdMap idm, uint32_t n_nodes,
                                                   uint32_t *__restrict__ dense) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t d = id_to_dense(idm, frags[i]);
        dense[i] = (d != 0xFFFFFFFFu && d < n_nodes) ? d + 1u : 0u;
    }
}

int dense_fragments(Plan &P, const uint64_t *frags, int64_t n, uint32_t *dense_out, cudaStream_t s) {
    if (n == 0) return BS_OK;
    DevBuf d_c2d;
    IdMap idm;
    BS_TRY(plan_idmap(P, d_c2d, &idm, s));
    BS_LAUNCH(k_dense_ids, (unsigned)std::min<size_t>(cdiv((size_t)n, 256), 148 * 32), 256, 0, s, frags, (size_t)n, idm,
              (uint32_t)P.block_nbase[P.blocks.size()], dense_out);
    BS_CUDA(cudaGetLastError());
    BS_CUDA(cudaStreamSynchronize(s));   // the id table is call-local scratch
    return BS_OK;
}

int plan_node_ids(Plan &P, uint64_t *out, long long *n_out, cudaStream_t s) {
    const size_t nb = P.blocks.size();
    const long long n = P.block_nbase[nb];
    if (n_out) *n_out = n;
    if (!out || n == 0) return BS_OK;
    std::vector<long long> h(2 * nb);
    for (size_t i = 0; i < nb; i++) {
        h[i] = P.block_nbase[i];
        h[nb + i] = P.blocks[i].block_id;
    }
    DevBuf d;
    BS_TRY(d.alloc(16 * nb, s));
    BS_CUDA(cudaMemcpyAsync(d.p, h.data(), 16 * nb, cudaMemcpyHostToDevice, s));
    BS_LAUNCH(k_node_ids, cdiv((size_t)n, 256), 256, 0, s, d.as<long long>(), d.as<long long>() + nb, (int)nb, P.nvox_block, n, out);
    BS_CUDA(cudaStreamSynchronize(s));   // h is a host-staged copy
    BS_CUDA(cudaGetLastError());
    return BS_OK;
}

}  // namespace bs
