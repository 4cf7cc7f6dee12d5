// Fragment-pair affinity means over an arbitrary neighbourhood: volara's AffAgglom as the blockwise mws pipeline drives it
// (post/watershed_mutex.py:144-154: AffAgglom(scores={"zyx_aff": neighborhood})).
//
// Per block: fragments and affinities of the read ROI (zero fill outside the array / the ROI); for every offset c of the
// neighbourhood and every voxel p with q = p + offset_c inside the read ROI, a pair of different non-zero fragments
// (frags[p], frags[q]) contributes affs[c][p] to the edge (min, max); the edge attribute is the mean over all contributions of
// all offsets.  An edge is written by the block that owns its smaller endpoint (same ownership rule as stage 2).
// Sums are exact integers (uint8: raw bytes; float32: rint(a * 2^38)) so the result does not depend on the order of the
// atomics; mean = float32(float64(sum) / 255 / count)  (float32 input: float32(ldexp(sum, -38) / count)).
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "geom.h"

namespace bs {

int plan_node_ids(Plan &P, uint64_t *out, long long *n_out, cudaStream_t s);

static constexpr uint32_t NONE32 = 0xFFFFFFFFu;
static constexpr unsigned long long EMPTY64 = 0xFFFFFFFFFFFFFFFFull;
static constexpr unsigned FULL = 0xFFFFFFFFu;

struct AaBlk {
    int ro[3], rs[3];                 // read ROI (volume coordinates)
    uint32_t own_first, own_count;    // dense node range of the block's own fragments
    uint32_t tbase, tcap;             // hash table slice
};

struct AaGeom {
    int C;
    int off[32][3];
    int volZ, volY, volX, wz0;        // affinity array (window) and its first z
    int roz, roy, rox, rsz, rsy, rsx; // fragments array: offset in volume coordinates, shape
};

__device__ __forceinline__ uint64_t aa_hash(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return k;
}

template <typename T>
__device__ __forceinline__ unsigned long long aa_fixed(T v);
template <>
__device__ __forceinline__ unsigned long long aa_fixed<uint8_t>(uint8_t v) {
    return v;
}
template <>
__device__ __forceinline__ unsigned long long aa_fixed<float>(float v) {
    return (unsigned long long)__double2ll_rn(ldexp((double)v, 38));
}

__global__ void __launch_bounds__(256) k_aa_dense(const uint64_t *__restrict__ frags, size_t n, IdMap idm, uint32_t *__restrict__ dense) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dense[i] = id_to_dense(idm, frags[i]) + 1u;   // NONE32 + 1 == 0
}

// one warp per row of the read ROI, lanes along x; lanes that hold the same pair add up before the table is touched
template <typename T>
__global__ void __launch_bounds__(256) k_aa_accumulate(const AaBlk *__restrict__ blks, const T *__restrict__ affs,
                                                       const uint32_t *__restrict__ frags, AaGeom G, unsigned long long *hkeys,
                                                       unsigned long long *hsum, uint32_t *hcnt, uint32_t *overflow) {
    const AaBlk &b = blks[blockIdx.y];
    const int RZ = b.rs[0], RY = b.rs[1], RX = b.rs[2];
    const size_t nvol = (size_t)G.volZ * G.volY * G.volX;
    const uint32_t tmask = b.tcap - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = blockIdx.x * 8 + warp; r < RZ * RY; r += gridDim.x * 8) {
        const int z = r / RY, y = r - z * RY;
        const int gz = b.ro[0] + z, gy = b.ro[1] + y;
        const int fz = gz - G.roz, fy = gy - G.roy;
        const bool row_in = fz >= 0 && fz < G.rsz && fy >= 0 && fy < G.rsy;
        for (int x0 = 0; x0 < RX; x0 += 32) {
            const int x = x0 + lane;
            const int gx = b.ro[2] + x, fx = gx - G.rox;
            const bool in = x < RX && row_in && fx >= 0 && fx < G.rsx;
            const uint32_t f1 = in ? frags[((size_t)fz * G.rsy + fy) * G.rsx + fx] : 0;
            for (int c = 0; c < G.C; c++) {
                const int qz = z + G.off[c][0], qy = y + G.off[c][1], qx = x + G.off[c][2];
                uint32_t f2 = 0;
                if (f1 && qz >= 0 && qz < RZ && qy >= 0 && qy < RY && qx >= 0 && qx < RX) {
                    const int hz = fz + G.off[c][0], hy = fy + G.off[c][1], hx = fx + G.off[c][2];
                    if (hz >= 0 && hz < G.rsz && hy >= 0 && hy < G.rsy && hx >= 0 && hx < G.rsx)
                        f2 = frags[((size_t)hz * G.rsy + hy) * G.rsx + hx];
                }
                const bool has = f2 != 0 && f2 != f1;
                const unsigned act = __ballot_sync(FULL, has);
                if (!has) continue;
                const uint32_t lo = min(f1, f2) - 1, hi = max(f1, f2) - 1;
                const unsigned long long key = ((unsigned long long)lo << 32) | hi;
                const unsigned long long a = aa_fixed<T>(affs[(size_t)c * nvol + ((size_t)(gz - G.wz0) * G.volY + gy) * G.volX + gx]);
                const unsigned peers = __match_any_sync(act, key);
                unsigned long long asum = 0;
                if constexpr (sizeof(T) == 1) {
                    asum = __reduce_add_sync(peers, (unsigned)a);
                } else {
                    unsigned rem = peers;
                    while (rem) {
                        const int src = __ffs(rem) - 1;
                        rem &= rem - 1;
                        asum += __shfl_sync(peers, a, src);
                    }
                }
                if (lane != __ffs(peers) - 1) continue;
                uint32_t slot = (uint32_t)aa_hash(key) & tmask, probes = 0;
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
                    atomicAdd(&hsum[(size_t)b.tbase + slot], asum);
                    atomicAdd(&hcnt[(size_t)b.tbase + slot], (uint32_t)__popc(peers));
                }
            }
        }
    }
}

// owned edges of all tables -> (dense pair key, mean) appended in any order (sorted afterwards)
template <bool U8>
__global__ void __launch_bounds__(256) k_aa_collect(const AaBlk *__restrict__ blks, const unsigned long long *__restrict__ hkeys,
                                                    const unsigned long long *__restrict__ hsum, const uint32_t *__restrict__ hcnt,
                                                    uint64_t *__restrict__ okeys, float *__restrict__ oscore, unsigned long long *__restrict__ n_out,
                                                    unsigned long long cap) {
    const AaBlk &b = blks[blockIdx.y];
    for (uint32_t sl = blockIdx.x * blockDim.x + threadIdx.x; sl < b.tcap; sl += gridDim.x * blockDim.x) {
        const size_t g = (size_t)b.tbase + sl;
        const unsigned long long key = hkeys[g];
        if (key == EMPTY64) continue;
        const uint32_t lo = (uint32_t)(key >> 32);
        if (lo < b.own_first || lo - b.own_first >= b.own_count) continue;
        const unsigned long long o = atomicAdd(n_out, 1ull);
        if (o >= cap) continue;
        double sum;
        if (U8)
            sum = __ddiv_rn((double)hsum[g], 255.0);
        else
            sum = ldexp((double)(long long)hsum[g], -38);
        okeys[o] = key;
        oscore[o] = __double2float_rn(__ddiv_rn(sum, (double)hcnt[g]));
    }
}

__global__ void k_aa_count(const AaBlk *__restrict__ blks, const unsigned long long *__restrict__ hkeys, unsigned long long *__restrict__ n_out) {
    const AaBlk &b = blks[blockIdx.y];
    unsigned long long n = 0;
    for (uint32_t sl = blockIdx.x * blockDim.x + threadIdx.x; sl < b.tcap; sl += gridDim.x * blockDim.x) {
        const unsigned long long key = hkeys[(size_t)b.tbase + sl];
        if (key == EMPTY64) continue;
        const uint32_t lo = (uint32_t)(key >> 32);
        if (lo >= b.own_first && lo - b.own_first < b.own_count) n++;
    }
    n = __reduce_add_sync(FULL, (unsigned)n);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(n_out, n);
}

__global__ void k_aa_emit(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, const float *__restrict__ score_in, size_t n,
                          const uint64_t *__restrict__ node_ids, uint64_t *__restrict__ eu, uint64_t *__restrict__ ev,
                          float *__restrict__ es) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint64_t k = keys[i];
        eu[i] = node_ids[(uint32_t)(k >> 32)];
        ev[i] = node_ids[(uint32_t)k];
        es[i] = score_in[vals[i]];
    }
}

__global__ void k_aa_iota(uint32_t *v, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) v[i] = (uint32_t)i;
}

static uint32_t aa_pow2(uint64_t v) {
    uint64_t p = 1;
    while (p < v) p <<= 1;
    return (uint32_t)std::min<uint64_t>(p, 1ull << 31);
}

template <typename T>
static int aff_agglom_impl(Plan &P, const T *affs, const uint64_t *frags, int C, const int32_t *offsets, int mult, bool *overflowed,
                           cudaStream_t s) {
    const bs_ws_config &cfg = P.cfg;
    const int nown = (int)P.owned.size();
    *overflowed = false;
    P.n_edges = 0;
    if (nown == 0) return BS_OK;
    BS_ARG(nown <= 65535, "bs_aff_agglom: too many owned blocks for one launch");
    const size_t nblocks = P.blocks.size();
    BS_ARG(P.block_nbase[nblocks] < (1LL << 32) - 2, "bs_aff_agglom: more than 2^32 fragments");
    std::vector<AaBlk> hb(nown);
    uint64_t tcur = 0;
    int maxrows = 1;
    for (int i = 0; i < nown; i++) {
        const Blk &b = P.blocks[P.owned[i]];
        AaBlk &d = hb[i];
        for (int k = 0; k < 3; k++) d.ro[k] = b.ro[k], d.rs[k] = b.rs[k];
        d.own_first = (uint32_t)P.block_nbase[P.owned[i]];
        d.own_count = (uint32_t)P.block_count[P.owned[i]];
        // distinct pairs: the block's own fragments plus the share of its neighbours' that reaches into the halo, each with a
        // few partners per offset; too small a table is detected and the call repeated with 4x the slots
        double halo = 0;
        for (int k = 0; k < 3; k++) halo += 2.0 * (double)cfg.context[k] / (double)std::max(1, cfg.block_size[k]);
        const double est = (double)P.block_count[P.owned[i]] * (1.0 + halo) + 64.0;
        d.tcap = aa_pow2((uint64_t)std::max(4096.0, 2.0 * (double)C * est * mult));
        d.tbase = (uint32_t)tcur;
        tcur += d.tcap;
        BS_ARG(tcur < (1ull << 32), "bs_aff_agglom: hash tables exceed 32-bit indexing");
        maxrows = std::max(maxrows, d.rs[0] * d.rs[1]);
    }
    const size_t Ttot = tcur;
    DevBuf d_blks, d_c2d, hkeys, hsum, hcnt, ovf, fdense, cnt;
    BS_TRY(d_blks.alloc(sizeof(AaBlk) * nown, s));
    BS_CUDA(cudaMemcpyAsync(d_blks.p, hb.data(), sizeof(AaBlk) * nown, cudaMemcpyHostToDevice, s));
    BS_TRY(hkeys.alloc_fill(8 * Ttot, 0xFF, s));
    BS_TRY(hsum.alloc_zero(8 * Ttot, s));
    BS_TRY(hcnt.alloc_zero(4 * Ttot, s));
    BS_TRY(ovf.alloc_zero(16, s));
    BS_TRY(cnt.alloc_zero(16, s));
    IdMap idm;
    BS_TRY(plan_idmap(P, d_c2d, &idm, s));
    AaGeom G;
    G.C = C;
    for (int c = 0; c < C; c++)
        for (int d = 0; d < 3; d++) G.off[c][d] = offsets[3 * c + d];
    const bool win = cfg.win_z > 0;
    G.volZ = win ? cfg.win_z : cfg.vol_shape[0], G.volY = cfg.vol_shape[1], G.volX = cfg.vol_shape[2];
    G.wz0 = win ? cfg.win_z0 : 0;
    G.roz = cfg.roi_offset[0] + (win ? cfg.win_z0 : 0), G.roy = cfg.roi_offset[1], G.rox = cfg.roi_offset[2];
    G.rsz = win ? cfg.win_z : cfg.roi_shape[0], G.rsy = cfg.roi_shape[1], G.rsx = cfg.roi_shape[2];
    const size_t nfr = (size_t)G.rsz * G.rsy * G.rsx;
    BS_TRY(fdense.alloc(4 * nfr, s));
    BS_LAUNCH(k_aa_dense, (unsigned)std::min<size_t>(cdiv(nfr, 256), 148 * 32), 256, 0, s, frags, nfr, idm, fdense.as<uint32_t>());
    dim3 gr((unsigned)std::min(std::max((maxrows + 15) / 16, 1), 4096), nown);
    BS_LAUNCH((k_aa_accumulate<T>), gr, 256, 0, s, d_blks.as<AaBlk>(), affs, fdense.as<uint32_t>(), G, hkeys.as<unsigned long long>(),
              hsum.as<unsigned long long>(), hcnt.as<uint32_t>(), ovf.as<uint32_t>());
    uint32_t maxcap = 0;
    for (auto &d : hb) maxcap = std::max(maxcap, d.tcap);
    dim3 gc(std::min<unsigned>(cdiv(maxcap, 256), 1024), nown);
    BS_LAUNCH(k_aa_count, gc, 256, 0, s, d_blks.as<AaBlk>(), hkeys.as<unsigned long long>(), cnt.as<unsigned long long>());
    unsigned long long h_n = 0;
    uint32_t h_ovf = 0;
    BS_CUDA(cudaMemcpyAsync(&h_n, cnt.p, 8, cudaMemcpyDeviceToHost, s));
    BS_CUDA(cudaMemcpyAsync(&h_ovf, ovf.p, 4, cudaMemcpyDeviceToHost, s));
    BS_CUDA(cudaStreamSynchronize(s));
    if (h_ovf) {
        *overflowed = true;
        return BS_OK;
    }
    const size_t E = (size_t)h_n;
    BS_ARG(E < (1ull << 32) - 1, "bs_aff_agglom: more than 2^32 edges");
    DevBuf keys, keys2, vals, vals2, score, ids;
    BS_TRY(keys.alloc(8 * (E + 1), s));
    BS_TRY(keys2.alloc(8 * (E + 1), s));
    BS_TRY(vals.alloc(4 * (E + 1), s));
    BS_TRY(vals2.alloc(4 * (E + 1), s));
    BS_TRY(score.alloc(4 * (E + 1), s));
    BS_CUDA(cudaMemsetAsync(cnt.p, 0, 8, s));
    if (sizeof(T) == 1)
        BS_LAUNCH((k_aa_collect<true>), gc, 256, 0, s, d_blks.as<AaBlk>(), hkeys.as<unsigned long long>(), hsum.as<unsigned long long>(),
                  hcnt.as<uint32_t>(), keys.as<uint64_t>(), score.as<float>(), cnt.as<unsigned long long>(), (unsigned long long)E);
    else
        BS_LAUNCH((k_aa_collect<false>), gc, 256, 0, s, d_blks.as<AaBlk>(), hkeys.as<unsigned long long>(), hsum.as<unsigned long long>(),
                  hcnt.as<uint32_t>(), keys.as<uint64_t>(), score.as<float>(), cnt.as<unsigned long long>(), (unsigned long long)E);
    long long nn = 0;
    BS_TRY(plan_node_ids(P, nullptr, &nn, s));
    BS_TRY(ids.alloc(8 * ((size_t)nn + 1), s));
    BS_TRY(plan_node_ids(P, ids.as<uint64_t>(), &nn, s));
    BS_TRY(P.edge_u.alloc_persistent(8 * (E + 1), s));
    BS_TRY(P.edge_v.alloc_persistent(8 * (E + 1), s));
    BS_TRY(P.edge_score.alloc_persistent(4 * (E + 1), s));
    if (E) {
        BS_LAUNCH(k_aa_iota, (unsigned)std::min<size_t>(cdiv(E, 256), 148 * 32), 256, 0, s, vals.as<uint32_t>(), E);
        BS_TRY(radix_sort_pairs(keys.as<uint64_t>(), vals.as<uint32_t>(), keys2.as<uint64_t>(), vals2.as<uint32_t>(), E, 0, 64, s));
        BS_LAUNCH(k_aa_emit, (unsigned)std::min<size_t>(cdiv(E, 256), 148 * 32), 256, 0, s, keys.as<uint64_t>(), vals.as<uint32_t>(),
                  score.as<float>(), E, ids.as<uint64_t>(), P.edge_u.as<uint64_t>(), P.edge_v.as<uint64_t>(), P.edge_score.as<float>());
    }
    P.n_edges = (long long)E;
    BS_CUDA(cudaStreamSynchronize(s));
    BS_CUDA(cudaGetLastError());
    return BS_OK;
}

int aff_agglom(Plan &P, const void *affs, const uint64_t *frags, int C, const int32_t *offsets, cudaStream_t s) {
    BS_ARG(C >= 1 && C <= 32, "bs_aff_agglom: 1..32 affinity channels");
    int mult = 1;
    for (int attempt = 0; attempt < 5; attempt++) {
        bool ovf = false;
        int rc = P.cfg.aff_dtype == BS_DTYPE_U8 ? aff_agglom_impl<uint8_t>(P, (const uint8_t *)affs, frags, C, offsets, mult, &ovf, s)
                                                : aff_agglom_impl<float>(P, (const float *)affs, frags, C, offsets, mult, &ovf, s);
        if (rc != BS_OK) return rc;
        if (!ovf) return BS_OK;
        mult *= 4;
    }
    set_error("bs_aff_agglom: hash table overflow after 5 attempts");
    return BS_ERR_OVERFLOW;
}

}  // namespace bs
