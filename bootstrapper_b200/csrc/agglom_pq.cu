// waterz agglomeration with the default (non-discretised) queue, as the single-shot ws path uses it
// (post/watershed.py:333-340: waterz.agglomerate(affs, thresholds, fragments, scoring_function) with
// discretize_queue = 0), for one region graph held in global memory.
//
// waterz pops the edge with the smallest (score, edge id) from a std::priority_queue; stale edges are re-scored
// and pushed back, deleted ones skipped, everything else merged.  The pop order is a closed
// total order, so the queue is emulated by a 32-ary min-heap of 64-bit keys (sortable score bits << 32 | edge):
// one warp owns the graph, a pop sifts down with one coalesced load of 32 children per level, a push climbs
// log32(E) levels.  The merge itself splices incidence lists and marks common neighbours by a generation stamp, on 32-bit indices.
// Thresholds are processed in ascending order; after each one the root of every node is written out, which is
// the segmentation waterz yields at that threshold.
//
// Scoring functions (post/watershed.py:232-244): OneMinus<MeanAffinity> on exact integer sums, or -- HIST --
// OneMinus<HistogramQuantileAffinity<Q, 256 bins>>: every edge owns a 256-bin histogram of its contact affinities (built
// by stage2.cu), a merge of two parallel edges adds the histograms (the warp adds 256 bins in 8 coalesced steps) and the
// score is 1 - (bin + 0.5) / 256 of the first bin whose running count reaches Q * sum / 100 + 1.
#include "agglom.cuh"

namespace bs {

static constexpr uint32_t NONE32 = 0xFFFFFFFFu;
static constexpr unsigned FULL = 0xFFFFFFFFu;
static constexpr int HD = 32;   // heap arity

__device__ __forceinline__ uint32_t score_bits(float s) {
    uint32_t u = __float_as_uint(s);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // total order of IEEE floats as unsigned
}

// initial scores + heap keys (edges are in creation order, the key's low word breaks score ties by edge id)
template <bool U8>
__global__ void k_pq_keys(const unsigned long long *__restrict__ esum, const uint32_t *__restrict__ ecnt, uint32_t E,
                          float *__restrict__ escore, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    float sc = edge_score<U8>(esum[e], ecnt[e]);
    escore[e] = sc;
    keys[e] = ((uint64_t)score_bits(sc) << 32) | e;
    vals[e] = e;
}

// quantile score of one histogram, computed by the whole warp (lane l owns bins 8l .. 8l+7)
__device__ __forceinline__ float hist_score(const uint32_t *h, int Q, int lane) {   // no __restrict__: the kernel updates histograms
    const uint4 p0 = *reinterpret_cast<const uint4 *>(h + lane * 8), p1 = *reinterpret_cast<const uint4 *>(h + lane * 8 + 4);
    const uint32_t b[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
    unsigned long long mine = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) mine += b[k];
    unsigned long long incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
    }
    const unsigned long long sum = __shfl_sync(FULL, incl, 31);
    const unsigned long long pivot = (unsigned long long)((long long)Q * (long long)sum / 100 + 1);
    const unsigned hit = __ballot_sync(FULL, incl >= pivot);
    int bin = 255;
    if (hit) {
        const int l = __ffs(hit) - 1;
        unsigned long long run = incl - mine;
        int mybin = 8 * lane + 7;
#pragma unroll
        for (int k = 7; k >= 0; k--) {
            unsigned long long upto = run;
#pragma unroll
            for (int j = 0; j <= k; j++) upto += b[j];
            if (upto >= pivot) mybin = 8 * lane + k;
        }
        bin = __shfl_sync(FULL, mybin, l);
    }
    const float q = __double2float_rn(((double)(float)bin + 0.5) / 256.0);   // undiscretize()
    return __double2float_rn(1.0 - (double)q);                               // OneMinus
}

__global__ void __launch_bounds__(256) k_pq_keys_hist(const uint32_t *__restrict__ hist, int Q, uint32_t E, float *__restrict__ escore,
                                                      uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const int lane = threadIdx.x & 31;
    const uint32_t e = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (e >= E) return;
    const float sc = hist_score(hist + (size_t)e * 256, Q, lane);
    if (lane == 0) {
        escore[e] = sc;
        keys[e] = ((uint64_t)score_bits(sc) << 32) | e;
        vals[e] = e;
    }
}

__device__ __forceinline__ uint32_t pq_find(uint32_t *ufp, uint32_t x) {
    for (;;) {
        uint32_t p = ufp[x];
        if (p == x) return x;
        uint32_t gp = ufp[p];
        if (gp == p) return p;
        ufp[x] = gp;
        x = gp;
    }
}

__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v, int lane, int &src) {
    // minimum over the warp and the lane that holds it
    uint64_t m = v;
    int s = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        uint64_t om = __shfl_xor_sync(FULL, m, o);
        int os = __shfl_xor_sync(FULL, s, o);
        if (om < m || (om == m && os < s)) {
            m = om;
            s = os;
        }
    }
    src = s;
    return m;
}

struct PqArgs {
    uint32_t E, N;
    const uint32_t *eu, *ev;           // compact endpoints
    unsigned long long *esum;
    uint32_t *ecnt;
    uint32_t *hist;                    // HIST: [E][256]
    int quantile;
    float *escore;
    uint32_t *etime;
    uint8_t *edead;
    uint32_t *anext, *ahead;           // half-edge lists
    uint32_t *ufp, *stamp, *mark, *markgen;
    uint64_t *heap;                    // E keys, sorted ascending on entry
    const float *thresholds;           // T ascending
    int T;
    uint32_t *roots;                   // [T][N]
    uint32_t *counters;                // pops, stale, dead, merges
    uint32_t *error;
    int keep_cheaper;
};

template <bool U8, bool HIST>
__global__ void __launch_bounds__(32) k_agglomerate_pq(PqArgs a) {
    const int lane = threadIdx.x;
    const uint32_t E = a.E, N = a.N;
    uint32_t *ufp = a.ufp, *stamp = a.stamp, *ahead = a.ahead, *anext = a.anext, *mark = a.mark, *markgen = a.markgen;
    uint64_t *heap = a.heap;
    for (uint32_t i = lane; i < N; i += 32) {
        ufp[i] = i;
        stamp[i] = 0;
        ahead[i] = NONE32;
        markgen[i] = 0;
    }
    for (uint32_t e = lane; e < E; e += 32) {
        a.etime[e] = 0;
        a.edead[e] = 0;
    }
    __syncwarp();
    // incidence lists (order is irrelevant): lanes that share a node chain their half-edges
    for (uint32_t e0 = 0; e0 < E; e0 += 32) {
        const uint32_t e = e0 + lane;
        const bool v = e < E;
        const unsigned act = __ballot_sync(FULL, v);
#pragma unroll
        for (int side = 0; side < 2; side++) {
            uint32_t node = 0;
            unsigned peers = 0;
            if (v) {
                node = side ? a.ev[e] : a.eu[e];
                peers = __match_any_sync(act, node);
                const unsigned higher = peers & ~((2u << lane) - 1u);
                anext[2 * e + side] = higher ? 2 * (e0 + (__ffs(higher) - 1)) + side : ahead[node];
            }
            __syncwarp();   // the old heads are read before any leader replaces them
            if (v && lane == __ffs(peers) - 1) ahead[node] = 2 * e + side;
            __syncwarp();
        }
    }

    uint32_t hn = E;   // heap size
    uint32_t n_pops = 0, n_stale = 0, n_dead = 0, nmerge = 0, clock = 0;
    bool fail = false;

    auto heap_pop = [&]() {   // removes the root (hn > 0)
        const uint64_t last = heap[--hn];
        if (hn == 0) return;
        uint32_t i = 0;
        for (;;) {
            const uint64_t c0 = (uint64_t)i * HD + 1;
            if (c0 >= hn) break;
            const uint64_t ci = c0 + lane;
            uint64_t ck = ci < hn ? heap[ci] : ~0ull;
            int src;
            const uint64_t mk = warp_min_u64(ck, lane, src);
            if (mk >= last) break;
            if (lane == 0) heap[i] = mk;
            i = (uint32_t)(c0 + src);
        }
        if (lane == 0) heap[i] = last;
        __syncwarp();
    };
    auto heap_push = [&](uint64_t key) {
        uint32_t i = hn++;
        while (i > 0) {
            const uint32_t par = (i - 1) / HD;
            const uint64_t pk = heap[par];
            if (pk <= key) break;
            if (lane == 0) heap[i] = pk;
            i = par;
        }
        if (lane == 0) heap[i] = key;
        __syncwarp();
    };
    auto walk = [&](uint32_t &head, uint32_t &tail, auto proc) {
        uint32_t h = head, prev = NONE32, guard = 0;
        for (;;) {
            int cnt = 0;
            uint32_t mineh = NONE32;
            while (h != NONE32 && cnt < 32) {
                if (++guard > 2 * E || h >= 2 * E) {   // corrupted list: stop instead of spinning
                    fail = true;
                    h = NONE32;
                    break;
                }
                const uint32_t nx = anext[h];
                if (a.edead[h >> 1]) {
                    if (prev == NONE32)
                        head = nx;
                    else if (lane == 0)
                        anext[prev] = nx;
                } else {
                    if (lane == cnt) mineh = h;
                    cnt++;
                    prev = h;
                }
                h = nx;
            }
            if (cnt == 0) break;
            __syncwarp();
            proc(mineh);
            __syncwarp();
            if (h == NONE32) break;
        }
        tail = prev;
    };

    for (int t = 0; t < a.T; t++) {
        const float thr = a.thresholds[t];
        while (hn > 0 && !fail) {
            const uint64_t top = heap[0];
            const uint32_t e = (uint32_t)top;
            const float sc = a.escore[e];
            if (sc >= thr) break;
            heap_pop();
            n_pops++;
            if (a.edead[e]) {
                n_dead++;
                continue;
            }
            const uint32_t ru = pq_find(ufp, a.eu[e]), rv = pq_find(ufp, a.ev[e]);
            const uint32_t te = a.etime[e];
            if (stamp[ru] > te || stamp[rv] > te) {
                // stale: re-score, push back
                float ns;
                if constexpr (HIST)
                    ns = hist_score(a.hist + (size_t)e * 256, a.quantile, lane);
                else
                    ns = edge_score<U8>(a.esum[e], a.ecnt[e]);
                if (lane == 0) {
                    a.escore[e] = ns;
                    a.etime[e] = clock;
                }
                __syncwarp();
                heap_push(((uint64_t)score_bits(ns) << 32) | e);
                n_stale++;
                continue;
            }
            // ---- merge: clusters ca < cb, ca survives (waterz mergeRegions)
            const uint32_t ca = min(ru, rv), cb = max(ru, rv);
            clock++;
            const uint32_t gen = clock;
            if (lane == 0) a.edead[e] = 1;
            __syncwarp();
            uint32_t head_b = ahead[cb], tail_b = NONE32;
            walk(head_b, tail_b, [&](uint32_t h) {
                if (h != NONE32) {
                    const uint32_t ne = h >> 1;
                    const uint32_t x1 = pq_find(ufp, a.eu[ne]), x2 = pq_find(ufp, a.ev[ne]);
                    const uint32_t x = x1 == cb ? x2 : x1;
                    mark[x] = ne;
                    markgen[x] = gen;
                }
            });
            uint32_t head_a = ahead[ca], tail_a = NONE32;
            walk(head_a, tail_a, [&](uint32_t h) {
                uint32_t dst = NONE32, src = NONE32;
                if (h != NONE32) {
                    const uint32_t ae = h >> 1;
                    const uint32_t x1 = pq_find(ufp, a.eu[ae]), x2 = pq_find(ufp, a.ev[ae]);
                    const uint32_t x = x1 == ca ? x2 : x1;
                    if (markgen[x] == gen) {
                        const uint32_t ne = mark[x];
                        if (!a.keep_cheaper || a.escore[ne] > a.escore[ae])
                            dst = ae, src = ne;
                        else
                            dst = ne, src = ae;
                        a.edead[src] = 1;
                        if constexpr (!HIST) {
                            a.esum[dst] += a.esum[src];
                            a.ecnt[dst] += a.ecnt[src];
                        }
                    }
                }
                if constexpr (HIST) {
                    // notifyEdgeMerge: the histograms add; one pair at a time, 256 bins across the warp
                    unsigned m = __ballot_sync(FULL, dst != NONE32);
                    while (m) {
                        const int l = __ffs(m) - 1;
                        m &= m - 1;
                        const uint32_t d = __shfl_sync(FULL, dst, l), sr = __shfl_sync(FULL, src, l);
                        uint32_t *hd = a.hist + (size_t)d * 256;
                        const uint32_t *hs = a.hist + (size_t)sr * 256;
#pragma unroll
                        for (int k = 0; k < 8; k++) hd[lane + 32 * k] += hs[lane + 32 * k];
                    }
                }
            });
            if (lane == 0) {
                if (head_b != NONE32) {
                    if (head_a == NONE32)
                        head_a = head_b;
                    else
                        anext[tail_a] = head_b;
                }
                ahead[ca] = head_a;
                ufp[cb] = ca;
                stamp[ca] = clock;
            }
            nmerge++;
            __syncwarp();
        }
        // the segmentation waterz yields at this threshold: every node's current root
        __syncwarp();
        for (uint32_t i = lane; i < N; i += 32) {
            uint32_t r = i;
            while (ufp[r] != r) r = ufp[r];
            a.roots[(size_t)t * N + i] = r;
        }
        __syncwarp();
    }
    if (lane == 0) {
        a.counters[0] = n_pops;
        a.counters[1] = n_stale;
        a.counters[2] = n_dead;
        a.counters[3] = nmerge;
        if (fail) atomicExch(a.error, 1u);
    }
}

// seg[t][i] = id of the root fragment of frags[i] at threshold t (ids that carry no edge keep their id)
struct PqRelabel {
    const uint32_t *roots;      // [T][Nc] compact roots
    const uint32_t *cscan;      // view node -> compact (exclusive scan of `used`)
    const uint8_t *used;
    const uint32_t *cmap;       // compact -> view node
    uint32_t Nc, nview;
    uint32_t dense0;            // dense number of view node 0
    long long block_id, nvox_block;
    uint64_t *seg[8];
    int T;
};

__global__ void __launch_bounds__(256) k_pq_relabel(const uint64_t *__restrict__ frags, size_t n, IdMap idm, PqRelabel r) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint64_t id = frags[i];
        uint32_t d = id_to_dense(idm, id);
        uint32_t v = d == NONE32 ? NONE32 : d - r.dense0;
        const bool hit = v < r.nview && r.used[v];
        const uint32_t c = hit ? r.cscan[v] : 0;
#pragma unroll
        for (int t = 0; t < 8; t++)
            if (t < r.T) {
                uint64_t out = id;
                if (hit) out = (uint64_t)(r.cmap[r.roots[(size_t)t * r.Nc + c]] + 1) + (uint64_t)r.block_id * (uint64_t)r.nvox_block;
                r.seg[t][i] = out;
            }
    }
}

__global__ void k_pq_cmap(const uint8_t *__restrict__ used, const uint32_t *__restrict__ cscan, uint32_t n, uint32_t *__restrict__ cmap) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && used[i]) cmap[cscan[i]] = i;
}

int agglom_pq_run(bool u8, uint32_t E, uint32_t Nc, const uint32_t *ceu, const uint32_t *cev, unsigned long long *esum,
                  uint32_t *ecnt, uint32_t *hist, int quantile, const float *thresholds_host, int T, int keep_cheaper,
                  uint32_t *roots_out, uint32_t *counters_host, cudaStream_t s) {
    DevBuf escore, etime, edead, anext, ahead, ufp, stamp, mark, markgen, keys, vals, keys2, vals2, thr, counters, err;
    BS_TRY(escore.alloc(4 * ((size_t)E + 1), s));
    BS_TRY(etime.alloc(4 * ((size_t)E + 1), s));
    BS_TRY(edead.alloc((size_t)E + 1, s));
    BS_TRY(anext.alloc(8 * ((size_t)E + 1), s));
    BS_TRY(ahead.alloc(4 * ((size_t)Nc + 1), s));
    BS_TRY(ufp.alloc(4 * ((size_t)Nc + 1), s));
    BS_TRY(stamp.alloc(4 * ((size_t)Nc + 1), s));
    BS_TRY(mark.alloc(4 * ((size_t)Nc + 1), s));
    BS_TRY(markgen.alloc(4 * ((size_t)Nc + 1), s));
    BS_TRY(keys.alloc(8 * ((size_t)E + 1), s));
    BS_TRY(vals.alloc(4 * ((size_t)E + 1), s));
    BS_TRY(keys2.alloc(8 * ((size_t)E + 1), s));
    BS_TRY(vals2.alloc(4 * ((size_t)E + 1), s));
    BS_TRY(thr.alloc(4 * (size_t)T, s));
    BS_TRY(counters.alloc_zero(32, s));
    BS_TRY(err.alloc_zero(16, s));
    BS_CUDA(cudaMemcpyAsync(thr.p, thresholds_host, 4 * (size_t)T, cudaMemcpyHostToDevice, s));
    if (E) {
        if (hist)
            BS_LAUNCH(k_pq_keys_hist, cdiv(E, 8), 256, 0, s, hist, quantile, E, escore.as<float>(), keys.as<uint64_t>(), vals.as<uint32_t>());
        else if (u8)
            BS_LAUNCH((k_pq_keys<true>), cdiv(E, 256), 256, 0, s, esum, ecnt, E, escore.as<float>(), keys.as<uint64_t>(),
                      vals.as<uint32_t>());
        else
            BS_LAUNCH((k_pq_keys<false>), cdiv(E, 256), 256, 0, s, esum, ecnt, E, escore.as<float>(), keys.as<uint64_t>(),
                      vals.as<uint32_t>());
        // a sorted array is a valid heap
        BS_TRY(radix_sort_pairs(keys.as<uint64_t>(), vals.as<uint32_t>(), keys2.as<uint64_t>(), vals2.as<uint32_t>(), E, 0, 64, s));
    }
    PqArgs a;
    a.E = E, a.N = Nc;
    a.eu = ceu, a.ev = cev;
    a.esum = esum, a.ecnt = ecnt;
    a.hist = hist, a.quantile = quantile;
    a.escore = escore.as<float>(), a.etime = etime.as<uint32_t>(), a.edead = edead.as<uint8_t>();
    a.anext = anext.as<uint32_t>(), a.ahead = ahead.as<uint32_t>();
    a.ufp = ufp.as<uint32_t>(), a.stamp = stamp.as<uint32_t>(), a.mark = mark.as<uint32_t>(), a.markgen = markgen.as<uint32_t>();
    a.heap = keys.as<uint64_t>();
    a.thresholds = thr.as<float>();
    a.T = T;
    a.roots = roots_out;
    a.counters = counters.as<uint32_t>();
    a.error = err.as<uint32_t>();
    a.keep_cheaper = keep_cheaper;
    if (hist)
        BS_LAUNCH((k_agglomerate_pq<true, true>), 1, 32, 0, s, a);
    else if (u8)
        BS_LAUNCH((k_agglomerate_pq<true, false>), 1, 32, 0, s, a);
    else
        BS_LAUNCH((k_agglomerate_pq<false, false>), 1, 32, 0, s, a);
    uint32_t h[5] = {0, 0, 0, 0, 0};
    BS_CUDA(cudaMemcpyAsync(h, counters.p, 16, cudaMemcpyDeviceToHost, s));
    BS_CUDA(cudaMemcpyAsync(h + 4, err.p, 4, cudaMemcpyDeviceToHost, s));
    BS_CUDA(cudaStreamSynchronize(s));
    if (h[4]) {
        set_error("agglomerate (priority queue): inconsistent incidence lists");
        return BS_ERR_STATE;
    }
    if (counters_host)
        for (int i = 0; i < 4; i++) counters_host[i] = h[i];
    BS_CUDA(cudaGetLastError());
    return BS_OK;
}

int agglom_pq_relabel(const uint64_t *frags, size_t n, IdMap idm, const uint32_t *roots, const uint32_t *cscan, const uint8_t *used,
                      uint32_t Nc, uint32_t nview, uint32_t dense0, long long block_id, long long nvox_block, int T,
                      uint64_t *const *segs, cudaStream_t s) {
    DevBuf cmap;
    BS_TRY(cmap.alloc(4 * ((size_t)Nc + 1), s));
    if (nview) BS_LAUNCH(k_pq_cmap, cdiv(nview, 256), 256, 0, s, used, cscan, nview, cmap.as<uint32_t>());
    for (int t0 = 0; t0 < T; t0 += 8) {
        PqRelabel r;
        r.roots = roots + (size_t)t0 * Nc;
        r.cscan = cscan, r.used = used, r.cmap = cmap.as<uint32_t>();
        r.Nc = Nc, r.nview = nview, r.dense0 = dense0;
        r.block_id = block_id, r.nvox_block = nvox_block;
        r.T = std::min(8, T - t0);
        for (int t = 0; t < 8; t++) r.seg[t] = t < r.T ? segs[t0 + t] : nullptr;
        unsigned grid = (unsigned)std::min<size_t>(cdiv(n, 256), 148 * 16 * 8);
        if (n) BS_LAUNCH(k_pq_relabel, grid, 256, 0, s, frags, n, idm, r);
    }
    BS_CUDA(cudaStreamSynchronize(s));
    BS_CUDA(cudaGetLastError());
    return BS_OK;
}

}  // namespace bs
