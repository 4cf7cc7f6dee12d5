// Stage 2: region adjacency graph + waterz agglomeration + merge-tree edge scores per owned block.
//
// Replaces WaterzAgglom.process_block (post/blockwise/waterz_agglom.py:106-170):
//   funlib.segment relabel (:116)             -> dense node numbers are a monotone map of the ids
//   waterz get_region_graph + MeanAffinity     -> k_rag_accumulate (hash table of (u,v) with exact
//                                                 integer affinity sums, counts and first-sight key)
//   waterz mergeUntil with BinQueue<256>       -> k_agglomerate (one warp per block, exact emulation of
//     (:131-139, thresholds [0, 1.0])             the FIFO bin queue incl. lazy stale re-scoring)
//   MergeTree replay + find_merges (:153-168)  -> merge tree built in k_agglomerate, k_lca per edge
//   write_graph ownership (:170)               -> edge kept by the block that owns node min(u, v)
#include <algorithm>

#include "agglom.cuh"

namespace bs {

static constexpr uint32_t NONE32 = 0xFFFFFFFFu;
static constexpr uint64_t EMPTY64 = 0xFFFFFFFFFFFFFFFFull;
static constexpr uint64_t TOMB64 = 0xFFFFFFFFFFFFFFFEull;
static constexpr unsigned FULL = 0xFFFFFFFFu;
extern int g_debug;
extern int g_agglom_version;   // 0 = shared-memory kernel when the block fits, 3 / 4 = force the global-slab kernels

struct S2Blk {
    long long block_id;
    int ro[3], rs[3];          // read ROI
    uint32_t own_first, own_count;   // dense node range of the block's own fragments
    uint32_t tbase, tcap;      // RAG hash table range (tcap power of two)
    uint32_t vbase, nv;        // view-local node range
    uint32_t nview;            // number of view ranges
    uint32_t view_first[27], view_count[27], view_prefix[27];   // ascending dense ranges of the 3x3x3 blocks
    long long view_block_id[27];
};

__device__ __forceinline__ uint64_t hash64(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return k;
}

// ------------------------------------------------------------------ RAG accumulation
// waterz get_region_graph: raster loop over the read ROI; for d in (z, y, x) pair p with
// p - e_d using affs[d][p]; ids 0 skipped.  Sums are exact integers (u8: raw bytes; f32: rint(a * 2^38)),
// DESIGN.md D2.  first = min over contributions of (raveled read-ROI index * 3 + d) = creation order.
template <typename T>
__device__ __forceinline__ unsigned long long aff_fixed(T v);
template <>
__device__ __forceinline__ unsigned long long aff_fixed<uint8_t>(uint8_t v) {
    return v;
}
template <>
__device__ __forceinline__ unsigned long long aff_fixed<float>(float v) {
    return (unsigned long long)__double2ll_rn(ldexp((double)v, 38));
}

// fragment ids of the task window -> dense node numbers + 1 (0: background or an id outside the plan), once per voxel:
// the RAG pass below then compares and keys 32-bit numbers instead of translating ids at every fragment boundary (in
// xy mode every voxel lies on one: the fragments of consecutive planes are different objects)
__global__ void __launch_bounds__(256) k_frags_dense(const uint64_t *__restrict__ frags, size_t n, IdMap idm,
                                                     uint32_t *__restrict__ dense) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dense[i] = id_to_dense(idm, frags[i]) + 1u;   // NONE32 + 1 == 0
}

template <typename T>
__global__ void __launch_bounds__(256) k_rag_accumulate(const S2Blk *__restrict__ blks, const T *__restrict__ affs,
                                                        const uint32_t *__restrict__ frags, int volZ, int volY,
                                                        int volX, int wz0, int roz, int roy, int rox, int rsz, int rsy, int rsx,
                                                        unsigned long long *hkeys, unsigned long long *hsum, uint32_t *hcnt,
                                                        uint32_t *hfirst, uint32_t *overflow) {
    const S2Blk &b = blks[blockIdx.y];
    const int RZ = b.rs[0], RY = b.rs[1], RX = b.rs[2];
    const size_t nvol = (size_t)volZ * volY * volX;
    const uint32_t tmask = b.tcap - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t plane = (size_t)rsy * rsx;
    // one warp per row of the read ROI: the row / plane predecessors are loaded once per 32 voxels and the
    // x predecessor comes from the neighbouring lane
    for (int r = blockIdx.x * 8 + warp; r < RZ * RY; r += gridDim.x * 8) {
        const int z = r / RY, y = r - z * RY;
        const int gz = b.ro[0] + z, gy = b.ro[1] + y;
        // fragments array covers the task ROI; outside: zero fill
        const int fz = gz - roz, fy = gy - roy;
        const bool row_in = fz >= 0 && fz < rsz && fy >= 0 && fy < rsy;
        const uint32_t *rowp = frags + (row_in ? ((size_t)fz * rsy + fy) * rsx : 0);
        const bool up_ok = row_in && y > 0 && fy > 0, zm_ok = row_in && z > 0 && fz > 0;
        uint32_t carry = 0;
      for (int x0 = 0; x0 < RX; x0 += 32) {
        const int x = x0 + lane;
        const int gx = b.ro[2] + x, fx = gx - rox;
        const bool in = x < RX && row_in && fx >= 0 && fx < rsx;
        const long long i = (long long)r * RX + x;
        const uint32_t f1 = in ? rowp[fx] : 0;
        uint32_t fxm = __shfl_up_sync(FULL, f1, 1);
        if (lane == 0) fxm = carry;
        carry = __shfl_sync(FULL, f1, 31);
        const uint32_t fym = (in && up_ok && f1) ? rowp[(long long)fx - rsx] : 0;
        const uint32_t fzm = (in && zm_ok && f1) ? rowp[(long long)fx - (long long)plane] : 0;
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const uint32_t f2 = f1 == 0 ? 0 : (d == 0 ? fzm : (d == 1 ? fym : fxm));
            const bool has = f2 != 0 && f2 != f1;
            unsigned act = __ballot_sync(FULL, has);
            if (has) {
                const uint32_t id1 = f1 - 1, id2 = f2 - 1;
                uint32_t lo = min(id1, id2), hi = max(id1, id2);
                unsigned long long key = ((unsigned long long)lo << 32) | hi;
                unsigned long long a = aff_fixed<T>(affs[(size_t)d * nvol + ((size_t)(gz - wz0) * volY + gy) * volX + gx]);
                uint32_t fk = (uint32_t)(i * 3 + d);
                unsigned peers = __match_any_sync(act, key);
                int leader = __ffs(peers) - 1;
                uint32_t fmin = __reduce_min_sync(peers, fk);
                unsigned long long asum;
                if constexpr (sizeof(T) == 1) {
                    asum = __reduce_add_sync(peers, (unsigned)a);
                } else {
                    // 64-bit segmented sum over the peer group
                    asum = 0;
                    unsigned rem = peers;
                    while (rem) {
                        int src = __ffs(rem) - 1;
                        rem &= rem - 1;
                        asum += __shfl_sync(peers, a, src);
                    }
                }
                if ((threadIdx.x & 31) == leader) {
                    uint32_t slot = (uint32_t)hash64(key) & tmask;
                    uint32_t probes = 0;
                    for (;;) {
                        unsigned long long *kp = &hkeys[(size_t)b.tbase + slot];
                        unsigned long long k = *((volatile unsigned long long *)kp);
                        if (k == EMPTY64) k = atomicCAS(kp, EMPTY64, key);
                        if (k == EMPTY64 || k == key) break;
                        slot = (slot + 1) & tmask;
                        if (++probes > tmask) {
                            atomicExch(overflow, 1u);
                            slot = NONE32;
                            break;
                        }
                    }
                    if (slot != NONE32) {
                        size_t s = (size_t)b.tbase + slot;
                        atomicAdd(&hsum[s], asum);
                        atomicAdd(&hcnt[s], (uint32_t)__popc(peers));
                        atomicMin(&hfirst[s], fmin);
                    }
                }
            }
        }
      }
    }
}

// ------------------------------------------------------------------ affinity histograms of the edges (single-shot path)
// waterz HistogramQuantileProvider::addAffinity (post/watershed.py:232-244): bin = min((int)(a * 256), 255) of the float32
// affinity the reference hands to waterz (uint8 input is normalised by / 255 first, post/watershed.py:254-257); bins below
// 0 (a shifted affinity < 0, undefined upstream) clamp to 0
template <typename T>
__device__ __forceinline__ int aff_bin(T v);
template <>
__device__ __forceinline__ int aff_bin<uint8_t>(uint8_t v) {
    return min(max((int)__fmul_rn(__fdiv_rn((float)v, 255.0f), 256.0f), 0), 255);
}
template <>
__device__ __forceinline__ int aff_bin<float>(float v) {
    return min(max((int)__fmul_rn(v, 256.0f), 0), 255);
}

__global__ void k_slot_edge(const uint32_t *__restrict__ svals, uint32_t E, uint32_t *__restrict__ slot_edge) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < E) slot_edge[svals[e]] = e;
}

// second pass over the contacts of block 0's read ROI: look the pair up in the (now complete) hash table, count the
// affinity's bin (or keep the largest bin, InitWithMax)
template <typename T>
__global__ void __launch_bounds__(256) k_rag_hist(const S2Blk *__restrict__ blks, const T *__restrict__ affs,
                                                  const uint32_t *__restrict__ frags, int volZ, int volY, int volX, int wz0, int roz,
                                                  int roy, int rox, int rsz, int rsy, int rsx,
                                                  const unsigned long long *__restrict__ hkeys, const uint32_t *__restrict__ slot_edge,
                                                  uint32_t *__restrict__ hist, int *__restrict__ emax) {
    const S2Blk &b = blks[0];
    const int RY = b.rs[1], RX = b.rs[2];
    const size_t n = (size_t)b.rs[0] * RY * RX, nvol = (size_t)volZ * volY * volX;
    const uint32_t tmask = b.tcap - 1;
    const int st[3] = {rsy * rsx, rsx, 1};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % RX), y = (int)((i / RX) % RY), z = (int)(i / ((size_t)RX * RY));
        const int gz = b.ro[0] + z, gy = b.ro[1] + y, gx = b.ro[2] + x;
        const int f[3] = {gz - roz, gy - roy, gx - rox};
        if (f[0] < 0 || f[0] >= rsz || f[1] < 0 || f[1] >= rsy || f[2] < 0 || f[2] >= rsx) continue;
        const size_t fi = ((size_t)f[0] * rsy + f[1]) * rsx + f[2];
        const uint32_t f1 = frags[fi];
        if (f1 == 0) continue;
        const int loc[3] = {z, y, x};
#pragma unroll
        for (int d = 0; d < 3; d++) {
            if (loc[d] == 0 || f[d] == 0) continue;
            const uint32_t f2 = frags[fi - st[d]];
            if (f2 == 0 || f2 == f1) continue;
            const uint32_t lo = min(f1, f2) - 1, hi = max(f1, f2) - 1;
            const unsigned long long key = ((unsigned long long)lo << 32) | hi;
            uint32_t slot = (uint32_t)hash64(key) & tmask;
            while (hkeys[(size_t)b.tbase + slot] != key) slot = (slot + 1) & tmask;   // present by construction
            const uint32_t e = slot_edge[(size_t)b.tbase + slot];
            const int bin = aff_bin<T>(affs[(size_t)d * nvol + ((size_t)(gz - wz0) * volY + gy) * volX + gx]);
            if (emax)
                atomicMax(&emax[e], bin);
            else
                atomicAdd(&hist[(size_t)e * 256 + bin], 1u);
        }
    }
}

__global__ void k_hist_initmax(const int *__restrict__ emax, uint32_t E, uint32_t *__restrict__ hist) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < E) hist[(size_t)e * 256 + emax[e]] = 1u;
}

__global__ void k_flag_keys(const unsigned long long *__restrict__ hkeys, uint8_t *__restrict__ flag, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = hkeys[i] != EMPTY64 ? 1 : 0;
}

__global__ void k_block_ebase(const S2Blk *__restrict__ blks, int nblk, const uint32_t *__restrict__ escan,
                              const uint32_t *__restrict__ etotal, uint32_t *__restrict__ ebase) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nblk) ebase[i] = escan[blks[i].tbase];
    if (i == nblk) ebase[i] = *etotal;
}

// sort keys: (block << 32 | first-sight key); value = table slot
__global__ void k_edge_sortkeys(const S2Blk *__restrict__ blks, const unsigned long long *__restrict__ hkeys,
                                const uint32_t *__restrict__ hfirst, const uint32_t *__restrict__ escan,
                                uint64_t *__restrict__ skeys, uint32_t *__restrict__ svals) {
    const S2Blk &b = blks[blockIdx.y];
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < b.tcap; s += gridDim.x * blockDim.x) {
        size_t g = (size_t)b.tbase + s;
        if (hkeys[g] != EMPTY64) {
            uint32_t o = escan[g];
            skeys[o] = ((uint64_t)blockIdx.y << 32) | hfirst[g];
            svals[o] = (uint32_t)g;
        }
    }
}

__device__ __forceinline__ uint32_t dense_to_local(const S2Blk &b, uint32_t dense) {
    for (uint32_t k = 0; k < b.nview; k++)
        if (dense >= b.view_first[k] && dense - b.view_first[k] < b.view_count[k])
            return b.view_prefix[k] + (dense - b.view_first[k]);
    return NONE32;
}

__device__ __forceinline__ uint64_t local_to_id(const S2Blk &b, uint32_t local, long long nvox_block) {
    for (uint32_t k = 0; k < b.nview; k++)
        if (local >= b.view_prefix[k] && local - b.view_prefix[k] < b.view_count[k])
            return (uint64_t)(local - b.view_prefix[k] + 1) + (uint64_t)b.view_block_id[k] * (uint64_t)nvox_block;
    return 0;
}

// gather sorted edges; endpoints become view-local node numbers; node degrees
__global__ void k_edge_gather(const S2Blk *__restrict__ blks, const uint32_t *__restrict__ ebase, int nblk,
                              const uint32_t *__restrict__ svals, const unsigned long long *__restrict__ hkeys,
                              const unsigned long long *__restrict__ hsum, const uint32_t *__restrict__ hcnt, uint32_t E,
                              uint32_t *__restrict__ eu, uint32_t *__restrict__ ev, unsigned long long *__restrict__ esum,
                              uint32_t *__restrict__ ecnt, uint32_t *__restrict__ eblk, uint32_t *__restrict__ deg) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    int lo = 0, hi = nblk - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (ebase[mid] <= e)
            lo = mid;
        else
            hi = mid - 1;
    }
    const S2Blk &b = blks[lo];
    size_t g = svals[e];
    unsigned long long key = hkeys[g];
    uint32_t u = dense_to_local(b, (uint32_t)(key >> 32)), v = dense_to_local(b, (uint32_t)key);
    eu[e] = u;
    ev[e] = v;
    esum[e] = hsum[g];
    ecnt[e] = hcnt[g];
    eblk[e] = lo;
    atomicAdd(&deg[b.vbase + u], 1u);
    atomicAdd(&deg[b.vbase + v], 1u);
}

// ---- block-compact node numbering: only fragments that carry at least one edge of the block take part in
// the agglomeration (the view of a block spans the fragments of all 27 neighbouring blocks)
__global__ void k_used_flag(const uint32_t *__restrict__ deg, uint8_t *__restrict__ used, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) used[i] = deg[i] ? 1 : 0;
}

__global__ void k_block_cbase(const S2Blk *__restrict__ blks, int nblk, const uint32_t *__restrict__ cscan, uint32_t nview,
                              const uint32_t *__restrict__ ctotal, uint32_t *__restrict__ cbase) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nblk) cbase[i] = blks[i].vbase < nview ? cscan[blks[i].vbase] : *ctotal;
    if (i == nblk) cbase[i] = *ctotal;
}

__global__ void k_compact_endpoints(const S2Blk *__restrict__ blks, const uint32_t *__restrict__ eblk,
                                    const uint32_t *__restrict__ eu, const uint32_t *__restrict__ ev, uint32_t E,
                                    const uint32_t *__restrict__ cscan, uint32_t *__restrict__ ceu, uint32_t *__restrict__ cev) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const uint32_t vb = blks[eblk[e]].vbase;
    const uint32_t c0 = cscan[vb];
    ceu[e] = cscan[vb + eu[e]] - c0;
    cev[e] = cscan[vb + ev[e]] - c0;
}

// ------------------------------------------------------------------ merge-tree score of every initial edge
// post/merge_tree.py:5-27: climb from the lower-level side until both sides meet; NaN if they never do.
__global__ void k_lca(const S2Blk *__restrict__ blks, const uint32_t *__restrict__ cbase, const uint32_t *__restrict__ eblk,
                      const uint32_t *__restrict__ eu, const uint32_t *__restrict__ ev, const uint32_t *__restrict__ ceu,
                      const uint32_t *__restrict__ cev, uint32_t E, const uint32_t *__restrict__ tparent,
                      const uint32_t *__restrict__ tlevel, const float *__restrict__ tscore, long long nvox_block,
                      uint64_t *__restrict__ out_u, uint64_t *__restrict__ out_v, float *__restrict__ out_s,
                      uint8_t *__restrict__ owned) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const S2Blk &b = blks[eblk[e]];
    const size_t tb = 2 * (size_t)cbase[eblk[e]];
    const uint32_t *tp = tparent + tb, *tl = tlevel + tb;
    const float *ts = tscore + tb;
    uint32_t u = ceu[e], v = cev[e];
    float score;
    for (;;) {
        if (u == v) {
            score = ts[u];
            break;
        }
        if (tl[u] > tl[v]) {
            uint32_t t = u;
            u = v;
            v = t;
        }
        uint32_t p = tp[u];
        if (p == NONE32) {
            score = __int_as_float(0x7fc00000);
            break;
        }
        u = p;
    }
    uint32_t lu = eu[e], lv = ev[e];
    out_u[e] = local_to_id(b, lu, nvox_block);
    out_v[e] = local_to_id(b, lv, nvox_block);
    out_s[e] = score;
    // write_graph(rag, block.write_roi): the edge is persisted by the block holding node min(u, v)
    // (its stored position, a centre of mass, always lies inside its own block's write ROI)
    uint32_t own_prefix = NONE32;
    for (uint32_t k = 0; k < b.nview; k++)
        if (b.view_first[k] == b.own_first && b.view_count[k] == b.own_count) own_prefix = b.view_prefix[k];
    uint32_t lo = min(lu, lv);
    owned[e] = (own_prefix != NONE32 && lo >= own_prefix && lo - own_prefix < b.own_count) ? 1 : 0;
}

__global__ void k_compact_edges(const uint8_t *__restrict__ owned, const uint32_t *__restrict__ oscan, uint32_t E,
                                const uint64_t *__restrict__ in_u, const uint64_t *__restrict__ in_v,
                                const float *__restrict__ in_s, uint64_t *__restrict__ out_u, uint64_t *__restrict__ out_v,
                                float *__restrict__ out_s) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E || !owned[e]) return;
    uint32_t o = oscan[e];
    out_u[o] = in_u[e];
    out_v[o] = in_v[e];
    out_s[o] = in_s[e];
}

// ------------------------------------------------------------------ host driver
static uint32_t next_pow2(uint64_t v) {
    uint64_t p = 1;
    while (p < v) p <<= 1;
    return (uint32_t)std::min<uint64_t>(p, 1ull << 31);
}

static int keep_debug2(Plan &P, const char *name, DevBuf &buf, int elem, long long count) {
    auto it = P.dbg.find(name);
    if (it != P.dbg.end()) {
        delete it->second;
        P.dbg.erase(it);
    }
    DevBuf *b = new DevBuf();
    b->swap(buf);
    P.dbg[name] = b;
    P.dbg_meta[name] = std::make_pair(elem, count);
    return BS_OK;
}

template <typename T>
static int stage2_impl(Plan &P, const void *affs, const uint64_t *frags, int table_mult, cudaStream_t s, bool *overflowed,
                       PqRequest *pq) {
    const bs_ws_config &cfg = P.cfg;
    const int nown = (int)P.owned.size();
    *overflowed = false;
    P.n_edges = 0;
    if (nown == 0) return BS_OK;
    BS_ARG(nown <= 65535, "stage2: too many owned blocks for one launch");
    // dense numbering of all blocks
    const size_t nblocks = P.blocks.size();
    BS_ARG(P.block_nbase[nblocks] < (1LL << 32) - 2, "stage2: more than 2^32 fragments");

    std::vector<S2Blk> hb(nown);
    uint64_t tcur = 0, vcur = 0;
    for (int i = 0; i < nown; i++) {
        const Blk &b = P.blocks[P.owned[i]];
        S2Blk &d = hb[i];
        d.block_id = b.block_id;
        for (int k = 0; k < 3; k++) d.ro[k] = b.ro[k], d.rs[k] = b.rs[k];
        d.own_first = (uint32_t)P.block_nbase[P.owned[i]];
        d.own_count = (uint32_t)P.block_count[P.owned[i]];
        // view ranges in ascending dense order (= ascending block id = ascending plan index)
        std::vector<int> nbs;
        for (int k = 0; k < 27; k++)
            if (b.nb[k] >= 0) nbs.push_back(b.nb[k]);
        std::sort(nbs.begin(), nbs.end());
        d.nview = 0;
        uint32_t pre = 0;
        double est = 0;
        for (int nbi : nbs) {
            uint32_t k = d.nview++;
            d.view_first[k] = (uint32_t)P.block_nbase[nbi];
            d.view_count[k] = (uint32_t)P.block_count[nbi];
            d.view_prefix[k] = pre;
            d.view_block_id[k] = P.blocks[nbi].block_id;
            pre += d.view_count[k];
            est += nbi == P.owned[i] ? (double)d.view_count[k] : 0.5 * d.view_count[k];
        }
        d.nv = pre;
        d.vbase = (uint32_t)vcur;
        vcur += pre;
        d.tcap = next_pow2((uint64_t)std::max(4096.0, 16.0 * est * table_mult));
        d.tbase = (uint32_t)tcur;
        tcur += d.tcap;
        BS_ARG(tcur < (1ull << 32) && vcur < (1ull << 31), "stage2: RAG tables exceed 32-bit indexing");
    }
    const size_t Ttot = tcur, Vtot = vcur;

    DevBuf d_blks, d_c2d, hkeys, hsum, hcnt, hfirst, ovf;
    BS_TRY(d_blks.alloc(sizeof(S2Blk) * nown, s));
    BS_CUDA(cudaMemcpyAsync(d_blks.p, hb.data(), sizeof(S2Blk) * nown, cudaMemcpyHostToDevice, s));
    BS_TRY(hkeys.alloc_fill(8 * Ttot, 0xFF, s));
    BS_TRY(hsum.alloc_zero(8 * Ttot, s));
    BS_TRY(hcnt.alloc_zero(4 * Ttot, s));
    BS_TRY(hfirst.alloc_fill(4 * Ttot, 0xFF, s));
    BS_TRY(ovf.alloc_zero(16, s));
    IdMap idm;
    BS_TRY(plan_idmap(P, d_c2d, &idm, s));
    const S2Blk *db = d_blks.as<S2Blk>();

    g_prof.mark("s2.rag", s);
    DevBuf fdense;
    long long maxread = 0;
    for (auto &d : hb) {
        long long rv = (long long)d.rs[0] * d.rs[1] * d.rs[2];
        BS_ARG(rv * 3 < (1LL << 32), "stage2: block read ROI too large for 32-bit first-sight keys");
        maxread = std::max(maxread, rv);
    }
    {
        int maxrows = 1;
        for (auto &d : hb) maxrows = std::max(maxrows, d.rs[0] * d.rs[1]);
        dim3 gr((unsigned)std::min(std::max((maxrows + 15) / 16, 1), 4096), nown);
        const size_t nfr = (size_t)(cfg.win_z > 0 ? cfg.win_z : cfg.roi_shape[0]) * cfg.roi_shape[1] * cfg.roi_shape[2];
        BS_TRY(fdense.alloc(4 * nfr, s));
        BS_LAUNCH(k_frags_dense, (unsigned)std::min<size_t>(cdiv(nfr, 256), 148 * 32), 256, 0, s, frags, nfr, idm, fdense.as<uint32_t>());
        BS_LAUNCH((k_rag_accumulate<T>), gr, 256, 0, s, db, (const T *)affs, fdense.as<uint32_t>(), cfg.win_z > 0 ? cfg.win_z : cfg.vol_shape[0],
                  cfg.vol_shape[1], cfg.vol_shape[2], cfg.win_z > 0 ? cfg.win_z0 : 0,
                  cfg.roi_offset[0] + (cfg.win_z > 0 ? cfg.win_z0 : 0), cfg.roi_offset[1], cfg.roi_offset[2],
                  cfg.win_z > 0 ? cfg.win_z : cfg.roi_shape[0], cfg.roi_shape[1], cfg.roi_shape[2], hkeys.as<unsigned long long>(), hsum.as<unsigned long long>(),
                  hcnt.as<uint32_t>(), hfirst.as<uint32_t>(), ovf.as<uint32_t>());
    }
    // ---- compaction + creation-order sort (host sync: number of edges)
    g_prof.mark("s2.edges", s);
    DevBuf flag, escan, tot, ebase;
    BS_TRY(flag.alloc(Ttot, s));
    BS_TRY(escan.alloc(4 * Ttot, s));
    BS_TRY(tot.alloc_zero(32, s));
    BS_TRY(ebase.alloc(4 * (nown + 1), s));
    BS_LAUNCH(k_flag_keys, cdiv(Ttot, 256), 256, 0, s, hkeys.as<unsigned long long>(), flag.as<uint8_t>(), Ttot);
    BS_TRY(scan_exclusive_u8(flag.as<uint8_t>(), escan.as<uint32_t>(), Ttot, tot.as<uint32_t>(), s));
    BS_LAUNCH(k_block_ebase, cdiv(nown + 1, 256), 256, 0, s, db, nown, escan.as<uint32_t>(), tot.as<uint32_t>(),
              ebase.as<uint32_t>());
    uint32_t h_tot[2] = {0, 0};
    BS_CUDA(cudaMemcpyAsync(&h_tot[0], tot.p, 4, cudaMemcpyDeviceToHost, s));
    BS_CUDA(cudaMemcpyAsync(&h_tot[1], ovf.p, 4, cudaMemcpyDeviceToHost, s));
    BS_CUDA(cudaStreamSynchronize(s));
    if (h_tot[1]) {
        *overflowed = true;
        return BS_OK;
    }
    const uint32_t E = h_tot[0];
    std::vector<uint32_t> h_ebase(nown + 1);
    BS_CUDA(cudaMemcpyAsync(h_ebase.data(), ebase.p, 4 * (nown + 1), cudaMemcpyDeviceToHost, s));
    BS_CUDA(cudaStreamSynchronize(s));

    DevBuf skeys, svals, skeys2, svals2;
    BS_TRY(skeys.alloc(8 * ((size_t)E + 1), s));
    BS_TRY(svals.alloc(4 * ((size_t)E + 1), s));
    BS_TRY(skeys2.alloc(8 * ((size_t)E + 1), s));
    BS_TRY(svals2.alloc(4 * ((size_t)E + 1), s));
    {
        uint32_t maxcap = 0;
        for (auto &d : hb) maxcap = std::max(maxcap, d.tcap);
        dim3 gr(std::min<unsigned>(cdiv(maxcap, 256), 1024), nown);
        BS_LAUNCH(k_edge_sortkeys, gr, 256, 0, s, db, hkeys.as<unsigned long long>(), hfirst.as<uint32_t>(),
                  escan.as<uint32_t>(), skeys.as<uint64_t>(), svals.as<uint32_t>());
    }
    int bbits = 0;
    while ((1 << bbits) < nown) bbits++;
    BS_TRY(radix_sort_pairs(skeys.as<uint64_t>(), svals.as<uint32_t>(), skeys2.as<uint64_t>(), svals2.as<uint32_t>(), E, 0,
                            32 + ((bbits + 7) / 8) * 8, s));
    skeys2.release();
    svals2.release();
    flag.release();
    escan.release();

    DevBuf eu, ev, esum, ecnt, eblk, deg;
    BS_TRY(eu.alloc(4 * ((size_t)E + 1), s));
    BS_TRY(ev.alloc(4 * ((size_t)E + 1), s));
    BS_TRY(esum.alloc(8 * ((size_t)E + 1), s));
    BS_TRY(ecnt.alloc(4 * ((size_t)E + 1), s));
    BS_TRY(eblk.alloc(4 * ((size_t)E + 1), s));
    BS_TRY(deg.alloc_zero(4 * (Vtot + 1), s));
    if (E)
        BS_LAUNCH(k_edge_gather, cdiv(E, 256), 256, 0, s, db, ebase.as<uint32_t>(), nown, svals.as<uint32_t>(),
                  hkeys.as<unsigned long long>(), hsum.as<unsigned long long>(), hcnt.as<uint32_t>(), E, eu.as<uint32_t>(),
                  ev.as<uint32_t>(), esum.as<unsigned long long>(), ecnt.as<uint32_t>(), eblk.as<uint32_t>(),
                  deg.as<uint32_t>());
    DevBuf ehist;
    if (pq && pq->quantile > 0 && E) {
        // single-shot path with a HistogramQuantileAffinity scoring function: per-edge affinity histograms
        BS_ARG(nown == 1, "histogram scoring needs a single-block plan");
        BS_ARG((size_t)E * 1024 <= ((size_t)64 << 30), "histogram scoring: more than 64 GB of edge histograms");
        DevBuf slot_edge, emax;
        BS_TRY(slot_edge.alloc(4 * Ttot, s));
        BS_TRY(ehist.alloc_zero((size_t)E * 1024, s));
        if (pq->initmax) BS_TRY(emax.alloc_zero(4 * ((size_t)E + 1), s));
        BS_LAUNCH(k_slot_edge, cdiv(E, 256), 256, 0, s, svals.as<uint32_t>(), E, slot_edge.as<uint32_t>());
        BS_LAUNCH((k_rag_hist<T>), (unsigned)std::min<size_t>(cdiv((size_t)maxread, 256), 148 * 32), 256, 0, s, db, (const T *)affs,
                  fdense.as<uint32_t>(), cfg.win_z > 0 ? cfg.win_z : cfg.vol_shape[0], cfg.vol_shape[1], cfg.vol_shape[2],
                  cfg.win_z > 0 ? cfg.win_z0 : 0, cfg.roi_offset[0] + (cfg.win_z > 0 ? cfg.win_z0 : 0), cfg.roi_offset[1],
                  cfg.roi_offset[2], cfg.win_z > 0 ? cfg.win_z : cfg.roi_shape[0], cfg.roi_shape[1], cfg.roi_shape[2],
                  hkeys.as<unsigned long long>(), slot_edge.as<uint32_t>(), ehist.as<uint32_t>(), pq->initmax ? emax.as<int>() : nullptr);
        if (pq->initmax) BS_LAUNCH(k_hist_initmax, cdiv(E, 256), 256, 0, s, emax.as<int>(), E, ehist.as<uint32_t>());
    }
    hkeys.release();
    hsum.release();
    hcnt.release();
    hfirst.release();
    skeys.release();
    svals.release();

    // ---- block-compact node numbers (host sync: compact node count per block)
    g_prof.mark("s2.adjacency", s);
    DevBuf used, cscan, cbase, ceu, cev;
    BS_TRY(used.alloc(Vtot + 1, s));
    BS_TRY(cscan.alloc(4 * (Vtot + 2), s));
    BS_TRY(cbase.alloc(4 * (nown + 1), s));
    BS_TRY(ceu.alloc(4 * ((size_t)E + 1), s));
    BS_TRY(cev.alloc(4 * ((size_t)E + 1), s));
    std::vector<uint32_t> h_cbase(nown + 1, 0);
    if (Vtot) {
        BS_LAUNCH(k_used_flag, cdiv(Vtot, 256), 256, 0, s, deg.as<uint32_t>(), used.as<uint8_t>(), Vtot);
        BS_TRY(scan_exclusive_u8(used.as<uint8_t>(), cscan.as<uint32_t>(), Vtot, tot.as<uint32_t>() + 2, s));
        BS_LAUNCH(k_block_cbase, cdiv(nown + 1, 256), 256, 0, s, db, nown, cscan.as<uint32_t>(), (uint32_t)Vtot,
                  tot.as<uint32_t>() + 2, cbase.as<uint32_t>());
        if (E)
            BS_LAUNCH(k_compact_endpoints, cdiv(E, 256), 256, 0, s, db, eblk.as<uint32_t>(), eu.as<uint32_t>(),
                      ev.as<uint32_t>(), E, cscan.as<uint32_t>(), ceu.as<uint32_t>(), cev.as<uint32_t>());
        BS_CUDA(cudaMemcpyAsync(h_cbase.data(), cbase.p, 4 * (nown + 1), cudaMemcpyDeviceToHost, s));
        BS_CUDA(cudaStreamSynchronize(s));
    }
    const size_t Ctot = h_cbase[nown];

    if (pq) {
        // ---- single-shot path: one region graph, non-discretised queue, a segmentation per threshold
        BS_ARG(nown == 1 && P.blocks.size() == 1, "simple_watershed needs a single-block plan (block = roi, no context)");
        BS_ARG(Ctot < (1ull << 31) && E < (1u << 31), "simple_watershed: region graph too large");
        g_prof.mark("s2.agglomerate", s);
        DevBuf roots;
        BS_TRY(roots.alloc(4 * ((size_t)pq->T * (Ctot + 1)), s));
        BS_TRY(agglom_pq_run(sizeof(T) == 1, E, (uint32_t)Ctot, ceu.as<uint32_t>(), cev.as<uint32_t>(), esum.as<unsigned long long>(),
                             ecnt.as<uint32_t>(), ehist.p ? ehist.as<uint32_t>() : nullptr, pq->quantile, pq->thresholds, pq->T, cfg.keep_cheaper,
                             roots.as<uint32_t>(), pq->counters, s));
        g_prof.mark("s2.relabel", s);
        const size_t nvox = (size_t)(cfg.win_z > 0 ? cfg.win_z : cfg.roi_shape[0]) * cfg.roi_shape[1] * cfg.roi_shape[2];
        BS_TRY(agglom_pq_relabel(frags, nvox, idm, roots.as<uint32_t>(), cscan.as<uint32_t>(), used.as<uint8_t>(), (uint32_t)Ctot,
                                 (uint32_t)Vtot, hb[0].view_first[0], hb[0].block_id, P.nvox_block, pq->T, pq->segs, s));
        return BS_OK;
    }

    // ---- which blocks fit the shared-memory kernel
    std::vector<AggBlk> ab(nown);
    std::vector<int> l_smem, l_glob;
    uint32_t Emax = 8, Nmax = 8;
    const bool u8 = sizeof(T) == 1;
    bool sum64 = !u8;
    for (int i = 0; i < nown; i++) {
        AggBlk &a = ab[i];
        a.ebase = h_ebase[i];
        a.E = h_ebase[i + 1] - h_ebase[i];
        a.vbase = h_cbase[i];
        a.nv = h_cbase[i + 1] - h_cbase[i];
        // u8 affinity sums of a whole block fit 32 bits when 3 * 255 * read voxels < 2^32
        if (u8 && 765.0 * hb[i].rs[0] * hb[i].rs[1] * hb[i].rs[2] >= 4294967295.0) sum64 = true;
    }
    // g_agglom_version: 0 = parallel-merge kernels (shared memory when the block fits, else a global slab),
    //                    3 = parallel-merge kernel on global slabs for every block,
    //                    4 = as 3 with every array in the slab (no shared-memory union-find / queue bins)
    {
        const size_t limit = 227 * 1024 - agglom_par_static_smem();
        for (int i = 0; i < nown; i++) {
            uint32_t Ec = (std::max<uint32_t>(ab[i].E, 8) + 7) & ~7u, Nc = (std::max<uint32_t>(ab[i].nv, 8) + 7) & ~7u;
            auto bytes = [&](uint32_t e_, uint32_t n_) { return agglom_par_bytes(e_, n_, sum64, 2); };
            bool fits = g_agglom_version == 0 && Ec <= 32760 && Nc <= 32760 && bytes(Ec, Nc) <= limit;
            if (fits && bytes(std::max(Emax, Ec), std::max(Nmax, Nc)) <= limit) {
                Emax = std::max(Emax, Ec);
                Nmax = std::max(Nmax, Nc);
                l_smem.push_back(i);
            } else {
                l_glob.push_back(i);
            }
        }
    }
    // blocks too large for shared memory: a global-memory slab each (32-bit indices, 64-bit sums)
    std::vector<unsigned long long> h_woff(l_glob.size() + 1, 0);
    uint32_t Nglob_max = 8;
    for (size_t k = 0; k < l_glob.size(); k++) {
        const AggBlk &a = ab[l_glob[k]];
        uint32_t Ec = (std::max<uint32_t>(a.E, 8) + 7) & ~7u, Nc = (std::max<uint32_t>(a.nv, 8) + 7) & ~7u;
        Nglob_max = std::max(Nglob_max, Nc);
        h_woff[k + 1] = h_woff[k] + agglom_par_bytes(Ec, Nc, true, 4);
    }
    const bool any_glob = !l_glob.empty();

    // ---- agglomeration
    g_prof.mark("s2.agglomerate", s);
    DevBuf d_ab, d_list, gwork, d_woff, tparent, tlevel, tscore, ha, hbb, hs, nmerges, counters, err;
    BS_TRY(d_ab.alloc(sizeof(AggBlk) * nown, s));
    BS_CUDA(cudaMemcpyAsync(d_ab.p, ab.data(), sizeof(AggBlk) * nown, cudaMemcpyHostToDevice, s));
    std::vector<int> h_list(l_smem);
    h_list.insert(h_list.end(), l_glob.begin(), l_glob.end());
    BS_TRY(d_list.alloc(sizeof(int) * (nown + 1), s));
    BS_CUDA(cudaMemcpyAsync(d_list.p, h_list.data(), sizeof(int) * nown, cudaMemcpyHostToDevice, s));
    if (any_glob) {
        BS_TRY(gwork.alloc((size_t)h_woff[l_glob.size()], s));
        BS_TRY(d_woff.alloc(8 * h_woff.size(), s));
        BS_CUDA(cudaMemcpyAsync(d_woff.p, h_woff.data(), 8 * h_woff.size(), cudaMemcpyHostToDevice, s));
    }
    BS_TRY(tparent.alloc(4 * (2 * Ctot + 2), s));
    BS_TRY(tlevel.alloc(4 * (2 * Ctot + 2), s));
    BS_TRY(tscore.alloc(4 * (2 * Ctot + 2), s));
    BS_TRY(ha.alloc(4 * (Ctot + 1), s));
    BS_TRY(hbb.alloc(4 * (Ctot + 1), s));
    BS_TRY(hs.alloc(4 * (Ctot + 1), s));
    BS_TRY(nmerges.alloc_zero(4 * nown, s));
    BS_TRY(counters.alloc_zero(24 * nown, s));
    BS_TRY(err.alloc_zero(16, s));
    AggArrays A;
    A.eu = ceu.as<uint32_t>(), A.ev = cev.as<uint32_t>(), A.ecnt = ecnt.as<uint32_t>();
    A.esum = esum.as<unsigned long long>();
    A.tparent = tparent.as<uint32_t>(), A.tlevel = tlevel.as<uint32_t>(), A.tscore = tscore.as<float>();
    A.ha = ha.as<uint32_t>(), A.hb = hbb.as<uint32_t>(), A.hs = hs.as<float>();
    A.nmerges = nmerges.as<uint32_t>();
    A.counters = counters.as<uint32_t>();
    A.error = err.as<uint32_t>();
    BS_ARG(cfg.queue_bins == 256, "stage2: only the BinQueue<256> agglomeration of the blockwise path is implemented");
    BS_TRY(agglom_par_launch(d_ab.as<AggBlk>(), d_list.as<int>(), (int)l_smem.size(), A, P.agg_threshold, cfg.keep_cheaper, u8, sum64, Emax,
                             Nmax, s));
    if (any_glob)
        BS_TRY(agglom_par_global_launch(d_ab.as<AggBlk>(), d_list.as<int>() + l_smem.size(), (int)l_glob.size(), A, P.agg_threshold,
                                        cfg.keep_cheaper, u8, gwork.as<unsigned char>(), d_woff.as<unsigned long long>(),
                                        Nglob_max, g_agglom_version != 4, s));

    // ---- merge-tree scores, ownership, output (host sync: number of owned edges)
    g_prof.mark("s2.lca", s);
    DevBuf ou, ov, os, owned, oscan;
    BS_TRY(ou.alloc(8 * ((size_t)E + 1), s));
    BS_TRY(ov.alloc(8 * ((size_t)E + 1), s));
    BS_TRY(os.alloc(4 * ((size_t)E + 1), s));
    BS_TRY(owned.alloc(((size_t)E + 1), s));
    BS_TRY(oscan.alloc(4 * ((size_t)E + 1), s));
    if (E) {
        BS_LAUNCH(k_lca, cdiv(E, 256), 256, 0, s, db, cbase.as<uint32_t>(), eblk.as<uint32_t>(), eu.as<uint32_t>(),
                  ev.as<uint32_t>(), ceu.as<uint32_t>(), cev.as<uint32_t>(), E, tparent.as<uint32_t>(), tlevel.as<uint32_t>(), tscore.as<float>(), P.nvox_block,
                  ou.as<uint64_t>(), ov.as<uint64_t>(), os.as<float>(), owned.as<uint8_t>());
        BS_TRY(scan_exclusive_u8(owned.as<uint8_t>(), oscan.as<uint32_t>(), E, tot.as<uint32_t>() + 1, s));
    }
    uint32_t h2[2] = {0, 0};
    BS_CUDA(cudaMemcpyAsync(&h2[0], tot.as<uint32_t>() + 1, 4, cudaMemcpyDeviceToHost, s));
    BS_CUDA(cudaMemcpyAsync(&h2[1], err.p, 4, cudaMemcpyDeviceToHost, s));
    BS_CUDA(cudaStreamSynchronize(s));
    if (h2[1]) {
        set_error("stage2: agglomeration workspace overflow (pair hash or queue pool)");
        return BS_ERR_OVERFLOW;
    }
    const uint32_t EO = E ? h2[0] : 0;
    BS_TRY(P.edge_u.alloc_persistent(8 * ((size_t)EO + 1), s));
    BS_TRY(P.edge_v.alloc_persistent(8 * ((size_t)EO + 1), s));
    BS_TRY(P.edge_score.alloc_persistent(4 * ((size_t)EO + 1), s));
    if (E)
        BS_LAUNCH(k_compact_edges, cdiv(E, 256), 256, 0, s, owned.as<uint8_t>(), oscan.as<uint32_t>(), E, ou.as<uint64_t>(),
                  ov.as<uint64_t>(), os.as<float>(), P.edge_u.as<uint64_t>(), P.edge_v.as<uint64_t>(), P.edge_score.as<float>());
    P.n_edges = EO;
    if (g_debug) {
        // everything the parity tests inspect: all per-block edges (creation order) with statistics and
        // scores, and the merge histories
        keep_debug2(P, "s2_ebase", ebase, 4, nown + 1);
        keep_debug2(P, "s2_eu", ou, 8, E);
        keep_debug2(P, "s2_ev", ov, 8, E);
        keep_debug2(P, "s2_escore", os, 4, E);
        keep_debug2(P, "s2_esum", esum, 8, E);
        keep_debug2(P, "s2_ecnt", ecnt, 4, E);
        keep_debug2(P, "s2_owned", owned, 1, E);
        keep_debug2(P, "s2_ha", ha, 4, Ctot);
        keep_debug2(P, "s2_hb", hbb, 4, Ctot);
        keep_debug2(P, "s2_hs", hs, 4, Ctot);
        keep_debug2(P, "s2_nmerges", nmerges, 4, nown);
        keep_debug2(P, "s2_counters", counters, 4, 6 * nown);
        std::vector<uint32_t> vb(nown);
        for (int i = 0; i < nown; i++) vb[i] = hb[i].vbase;
        DevBuf dvb;
        BS_TRY(dvb.alloc(4 * nown, s));
        BS_CUDA(cudaMemcpyAsync(dvb.p, vb.data(), 4 * nown, cudaMemcpyHostToDevice, s));
        BS_CUDA(cudaStreamSynchronize(s));
        keep_debug2(P, "s2_vbase", dvb, 4, nown);
    }
    BS_CUDA(cudaGetLastError());
    return BS_OK;
}

int stage2_run(Plan &P, const void *affs, const uint64_t *frags, cudaStream_t s, PqRequest *pq) {
    g_prof.reset();
    int mult = 1;
    for (int attempt = 0; attempt < 4; attempt++) {
        bool ovf = false;
        BS_TRY(g_arena.begin(!g_debug));
        int rc = P.cfg.aff_dtype == BS_DTYPE_U8 ? stage2_impl<uint8_t>(P, affs, frags, mult, s, &ovf, pq)
                                                : stage2_impl<float>(P, affs, frags, mult, s, &ovf, pq);
        g_arena.end();
        if (rc != BS_OK) return rc;
        if (!ovf) {
            g_prof.finish(s);
            return BS_OK;
        }
        mult *= 4;
    }
    set_error("stage2: RAG hash table overflow after 4 attempts");
    return BS_ERR_OVERFLOW;
}

}  // namespace bs
