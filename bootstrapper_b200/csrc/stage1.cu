// Stage 1: affinities -> supervoxel fragments for every owned block (sm_100a).
//
// Replaces WatershedFrags.process_block (post/blockwise/watershed_frags.py:196-246) and what it
// calls: watershed_from_affinities / watershed_from_boundary_distance (post/ws.py:8-112),
// filter_avg_fragments (watershed_frags.py:148-156), remove_small_objects (:188-192), crop +
// skimage.measure.label + global ids (:216-224), node statistics (:230-246).
//
// Kernel chain per batch of tiles (a tile = one z-slice of a block's read ROI in xy mode, or the
// whole read ROI in 3-D mode); every array is tile-concatenated, tile-local indices are u32:
//   k_mask_rowdist   boundary mask (integer test for u8, replayed f32 ops for f32) + distance to the
//                    nearest background pixel along x (warp per row, ballot bit rows in smem)
//   k_coldist/zdist  exact squared EDT by pruned lower-envelope search along y (and z)
//   k_maxfilt        separable maximum filter, size = min_seed_distance, scipy 'reflect' borders
//   k_seed_*         seeds = (maxfilter == d2) & mask, 4/6-connected components by lock-free
//                    min-root union-find
//   k_hist/k_levels  per-tile histogram of d2 -> dense priority levels + FIFO segment per level
//   k_flood          exact emulation of skimage's (value, age) priority flood: one warp per tile,
//                    32 queue items per step, neighbours claimed with atomicMin(rank*8+slot),
//                    step truncated at the first item that pushes a higher-priority pixel,
//                    order-preserving multi-level append
//   k_fragstats      per-fragment affinity sum + voxel count (warp-aggregated atomics)
//   k_crop_*         keep/drop decision, 8/26-connected relabel of the cropped write ROI
//   k_finalize       raster-order ids (+ block_id * prod(block_size)), uint64 output, node statistics
#include <cuda.h>

#include <atomic>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <type_traits>

#include "geom.h"

namespace bs {

static constexpr uint32_t UNLAB = 0xFFFFFFFFu;   // in mask, not labelled yet
static constexpr uint32_t CLAIM = 0x80000000u;   // claim keys live in [CLAIM, UNLAB)
static constexpr uint32_t NONE32 = 0xFFFFFFFFu;
static constexpr uint16_t GINF = 0xFFFF;
static constexpr uint32_t DBIG = 0x3FFFFFFFu;
static constexpr int MAXW = 4096;
// pixels a 256-thread CTA of the per-pixel kernels walks through: enough work to amortise the tile set-up, enough CTAs
// (tiles x pixels / PIX_PER_CTA) to fill 148 SMs on small batches
// true if this device's bit was already set in `mask` (and sets it)
static bool dev_once(std::atomic<unsigned long long> &mask) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
    const unsigned long long bit = 1ull << dev;
    return (mask.fetch_or(bit) & bit) != 0;
}

static const long long PIX_PER_CTA = getenv("BS_PIX_PER_CTA") ? atoll(getenv("BS_PIX_PER_CTA")) : 4096;
static constexpr unsigned FULL = 0xFFFFFFFFu;

// rank of position w in a bitmap whose per-word popcounts were scanned into wscan (exclusive)
__device__ __forceinline__ uint32_t bit_rank(const uint32_t *__restrict__ bits, const uint32_t *__restrict__ wscan, uint32_t w) {
    return wscan[w >> 5] + __popc(bits[w >> 5] & ((1u << (w & 31)) - 1u));
}
__device__ __forceinline__ bool bit_test(const uint32_t *__restrict__ bits, uint32_t w) { return (bits[w >> 5] >> (w & 31)) & 1u; }

struct AffView {
    const void *p;
    const uint8_t *mask;
    int C, Z, Y, X;   // global volume shape (zero fill outside)
    int z0, Zw;       // the array holds the global planes [z0, z0 + Zw) (multi-GPU slab window; else 0, Z)
};

// ------------------------------------------------------------------ affinity access
template <typename T>
struct AffOps;
template <>
struct AffOps<uint8_t> {
    // post/ws.py:64,77 on uint8/255 in float64 reduces exactly to a_y + a_x > 255 (2-D) and
    // post/ws.py:100 to a_z + a_y + a_x > 382 (3-D)  (exhaustively probed)
    static __device__ __forceinline__ bool boundary(const uint8_t *a, size_t n, size_t i, int ndim) {
        int ay = a[n + i], ax = a[2 * n + i];
        if (ndim == 2) return ay + ax > 255;
        return (int)a[i] + ay + ax > 382;
    }
    typedef unsigned long long acc_t;
    static __device__ __forceinline__ acc_t value(const uint8_t *a, size_t n, size_t i) {
        return (acc_t)a[i] + a[n + i] + a[2 * n + i];
    }
};
template <>
struct AffOps<float> {
    static __device__ __forceinline__ bool boundary(const float *a, size_t n, size_t i, int ndim) {
        float ay = a[n + i], ax = a[2 * n + i];
        if (ndim == 2) return __fmul_rn(0.5f, __fadd_rn(ax, ay)) > 0.5f;
        return __fdiv_rn(__fadd_rn(__fadd_rn(a[i], ay), ax), 3.0f) > 0.5f;
    }
    typedef double acc_t;
    static __device__ __forceinline__ acc_t value(const float *a, size_t n, size_t i) {
        return (double)__fdiv_rn(__fadd_rn(__fadd_rn(a[i], a[n + i]), a[2 * n + i]), 3.0f);
    }
};

// ---- optional shifts (watershed_frags.py:118-146): the watershed sees affs_data + shift with
//   shift = bias[c] - seed_eps * EDT(seeds == 0)      (noise_eps / sigma are not supported)
// replayed in the dtype numpy uses: float64 for uint8 input (normalised /255 in float64, watershed_frags.py:198-200),
// float32 for float32 input (the shift array is zeros_like(affs), in-place updates round to float32).
struct PreRef {          // the seed_eps pre-pass tile (= read ROI of the block) a main tile looks its distances up in
    long long base;
    int oz, oy, ox, H, W, D;
    long long block_id;
};
// seeded stand-in for numpy's unseeded randn (bsnative.h): unit-variance sum of four 16-bit uniforms
__device__ __forceinline__ uint64_t nz_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t nz_mixw(uint64_t x, long long w) { return nz_splitmix64(x ^ ((uint64_t)w * 0x9E3779B97F4A7C15ull)); }
__device__ __forceinline__ double seeded_normal(uint64_t seed, long long block_id, int c, long long idx) {
    const uint64_t h = nz_mixw(nz_mixw(nz_mixw(nz_mixw(seed, block_id), c), idx), 13);
    const long long sum = (long long)(h & 0xFFFF) + (long long)((h >> 16) & 0xFFFF) + (long long)((h >> 32) & 0xFFFF) + (long long)(h >> 48);
    return __dmul_rn((double)(sum - 131070), 1.7320508075688772 / 65536.0);
}
struct ShiftView {
    int has_bias, has_eps, has_sigma, has_noise;
    double noise_eps;
    unsigned long long noise_seed;
    double bias[3];
    double eps;
    const uint32_t *D2;      // squared distance to the nearest seed, per pre-pass tile pixel
    const void *G;           // gaussian-filtered affinities [3][gstride] in the pre-pass tile layout (double / float)
    size_t gstride;
    const PreRef *pre;       // per main tile
};

// inmask: voxel inside the volume and not masked out (else the normalised affinity is 0.0, to which the shift is added)
template <typename T>
__device__ __forceinline__ bool boundary_shifted(const T *a, size_t n, size_t i, bool inmask, int ndim, const ShiftView &S,
                                                 double dist, long long q, long long block_id, long long ridx, long long rvox);
template <>
__device__ __forceinline__ bool boundary_shifted<uint8_t>(const uint8_t *a, size_t n, size_t i, bool inmask, int ndim,
                                                          const ShiftView &S, double dist, long long q, long long block_id,
                                                          long long ridx, long long rvox) {
    double v[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        if (c == 0 && ndim == 2) continue;
        double x = inmask ? __ddiv_rn((double)a[(size_t)c * n + i], 255.0) : 0.0;
        double sh = 0.0;
        if (S.has_noise) sh = __dmul_rn(seeded_normal(S.noise_seed, block_id, c, (long long)c * rvox + ridx), S.noise_eps);   // zeros += randn * eps
        if (S.has_sigma) sh = __dadd_rn(sh, __dsub_rn(((const double *)S.G)[(size_t)c * S.gstride + q], x));   // shift += gaussian - affs
        if (S.has_bias) sh = __dadd_rn(sh, S.bias[c]);                 // shift += bias
        if (S.has_eps) sh = __dsub_rn(sh, __dmul_rn(S.eps, dist));     // shift -= seed_eps * D
        v[c] = __dadd_rn(x, sh);
    }
    if (ndim == 2) return __dmul_rn(0.5, __dadd_rn(v[2], v[1])) > 0.5;                // post/ws.py:64,77
    return __ddiv_rn(__dadd_rn(__dadd_rn(v[0], v[1]), v[2]), 3.0) > 0.5;              // post/ws.py:100
}
template <>
__device__ __forceinline__ bool boundary_shifted<float>(const float *a, size_t n, size_t i, bool inmask, int ndim,
                                                        const ShiftView &S, double dist, long long q, long long block_id,
                                                        long long ridx, long long rvox) {
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        if (c == 0 && ndim == 2) continue;
        float x = inmask ? a[(size_t)c * n + i] : 0.0f;
        float sh = 0.0f;
        // float32 shift array += float64 noise: float32(0 + n * eps)
        if (S.has_noise) sh = __double2float_rn(__dmul_rn(seeded_normal(S.noise_seed, block_id, c, (long long)c * rvox + ridx), S.noise_eps));
        if (S.has_sigma) sh = __fadd_rn(sh, __fsub_rn(((const float *)S.G)[(size_t)c * S.gstride + q], x));   // float32 arrays
        if (S.has_bias) sh = __double2float_rn(__dadd_rn((double)sh, S.bias[c]));                        // float32(shift + bias)
        if (S.has_eps) sh = __double2float_rn(__dsub_rn((double)sh, __dmul_rn(S.eps, dist)));            // float32(shift - eps * D)
        v[c] = __fadd_rn(x, sh);
    }
    if (ndim == 2) return __fmul_rn(0.5f, __fadd_rn(v[2], v[1])) > 0.5f;
    return __fdiv_rn(__fadd_rn(__fadd_rn(v[0], v[1]), v[2]), 3.0f) > 0.5f;
}

// ------------------------------------------------------------------ mask + row distance
// MODE 0: boundary mask from the affinities; 1: from the shifted affinities; 2: "not a seed" (pred[p] == NONE32),
// the input of the seed-distance transform EDT(seeds == 0).
template <typename T, int MODE>
__global__ void __launch_bounds__(256) k_mask_rowdist(const Tile *__restrict__ tiles, AffView A, ShiftView S,
                                                      const uint32_t *__restrict__ pred, uint8_t *__restrict__ msk,
                                                      uint16_t *__restrict__ g, uint32_t *__restrict__ tileflags) {
    __shared__ uint32_t bits[8][MAXW / 32];
    const Tile t = tiles[blockIdx.y];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rows = t.D * t.H, W = t.W, nw = (W + 31) >> 5;
    const size_t nvol = (size_t)A.Zw * A.Y * A.X;
    const T *a = (const T *)A.p;
    bool anybg = false;
    for (int r = blockIdx.x * 8 + warp; r < rows; r += gridDim.x * 8) {
        int z = r / t.H, y = r - z * t.H;
        int gz = t.gz + z, gy = t.gy + y;
        bool rowin = gz >= 0 && gz < A.Z && gy >= 0 && gy < A.Y && gz >= A.z0 && gz < A.z0 + A.Zw;
        size_t rowoff = rowin ? ((size_t)(gz - A.z0) * A.Y + gy) * A.X : 0;
        long long pbase = t.base + (long long)r * W;
        for (int c = 0; c < nw; c++) {
            int x = c * 32 + lane;
            bool m = false;
            if (MODE == 0) {
                if (x < W && rowin) {
                    int gx = t.gx + x;
                    if (gx >= 0 && gx < A.X) {
                        size_t i = rowoff + gx;
                        if (!A.mask || A.mask[i] > 0) m = AffOps<T>::boundary(a, nvol, i, t.ndim);
                    }
                }
            } else if (MODE == 1) {
                if (x < W) {
                    int gx = t.gx + x;
                    bool inside = rowin && gx >= 0 && gx < A.X;
                    size_t i = inside ? rowoff + gx : 0;
                    bool inmask = inside && (!A.mask || A.mask[i] > 0);
                    double dist = 0.0;
                    long long q = 0, ridx = 0, rvox = 0, bid = 0;
                    if (S.has_eps || S.has_sigma || S.has_noise) {
                        const PreRef pr = S.pre[blockIdx.y];
                        ridx = ((long long)(gz - pr.oz) * pr.H + (gy - pr.oy)) * pr.W + (gx - pr.ox);   // raveled read-ROI voxel
                        rvox = (long long)pr.D * pr.H * pr.W;
                        bid = pr.block_id;
                        q = pr.base + ridx;
                        if (S.has_eps) dist = __dsqrt_rn((double)S.D2[q]);
                    }
                    m = boundary_shifted<T>(a, nvol, i, inmask, t.ndim, S, dist, q, bid, ridx, rvox);
                }
            } else {
                if (x < W) m = pred[pbase + x] == NONE32;
            }
            unsigned b = __ballot_sync(FULL, m);
            int rem = W - c * 32;
            unsigned valid = rem >= 32 ? FULL : ((1u << rem) - 1u);
            unsigned bg = ~b & valid;
            if (lane == 0) bits[warp][c] = bg;
            anybg |= (bg != 0);
            if (MODE != 2 && x < W) msk[pbase + x] = m ? 1 : 0;
        }
        __syncwarp();
        for (int c = 0; c < nw; c++) {
            int x = c * 32 + lane;
            if (x < W) {
                int w = x >> 5, bpos = x & 31;
                uint32_t word = bits[warp][w];
                uint32_t dist;
                if ((word >> bpos) & 1u) {
                    dist = 0;
                } else {
                    uint32_t dl = GINF, dr = GINF;
                    uint32_t wl = word & ((1u << bpos) - 1u);
                    for (int ww = w;;) {
                        if (wl) {
                            dl = x - (ww * 32 + 31 - __clz(wl));
                            break;
                        }
                        if (--ww < 0) break;
                        wl = bits[warp][ww];
                    }
                    uint32_t wr = word & ~((2u << bpos) - 1u);
                    for (int ww = w;;) {
                        if (wr) {
                            dr = (ww * 32 + __ffs(wr) - 1) - x;
                            break;
                        }
                        if (++ww >= nw) break;
                        wr = bits[warp][ww];
                    }
                    dist = min(dl, dr);
                }
                g[pbase + x] = (uint16_t)dist;
            }
        }
        __syncwarp();
    }
    if (MODE != 2 && anybg && lane == 0) atomicOr(&tileflags[blockIdx.y], 1u);
}

// ------------------------------------------------------------------ exact squared EDT, y and z passes
__device__ __forceinline__ uint32_t warp_max_u32(uint32_t v) { return __reduce_max_sync(FULL, v); }

// in-plane pass: out = min_y' g(y',x)^2 + (y-y')^2 ; DBIG if the slice has no background.
// final2d: the tile is a 2-D array -> apply scipy's all-foreground rule and record the tile maximum.
__global__ void __launch_bounds__(256) k_coldist(const Tile *__restrict__ tiles, const uint16_t *__restrict__ g,
                                                 uint32_t *__restrict__ out, uint32_t *__restrict__ tilemax) {
    const Tile t = tiles[blockIdx.y];
    const int W = t.W, H = t.H;
    const long long npix = (long long)t.D * H * W;
    const bool final2d = (t.ndim == 2);
    uint32_t mymax = 0;
    for (long long i0 = (long long)blockIdx.x * blockDim.x; i0 < npix; i0 += (long long)gridDim.x * blockDim.x) {
        long long i = i0 + threadIdx.x;
        uint32_t best = 0;
        if (i < npix) {
            int x, y, zq;
            unravel3(i, W, H, x, y, zq);
            const uint16_t *gp = g + t.base + i;
            uint32_t g0 = gp[0];
            best = g0 == GINF ? DBIG : g0 * g0;
            for (int dy = 1;; dy++) {
                uint32_t dd = (uint32_t)dy * dy;
                if (dd >= best) break;
                bool any = false;
                if (y - dy >= 0) {
                    any = true;
                    uint32_t v = gp[-(long long)dy * W];
                    if (v != GINF) best = min(best, v * v + dd);
                }
                if (y + dy < H) {
                    any = true;
                    uint32_t v = gp[(long long)dy * W];
                    if (v != GINF) best = min(best, v * v + dd);
                }
                if (!any) break;
            }
            if (final2d && best == DBIG) best = (uint32_t)(y + 1) * (y + 1) + (uint32_t)x * x;
            out[t.base + i] = best;
            if (final2d) mymax = max(mymax, best);
        }
    }
    if (final2d) {
        mymax = warp_max_u32(mymax);
        if ((threadIdx.x & 31) == 0 && mymax) atomicMax(&tilemax[blockIdx.y], mymax);
    }
}

__global__ void __launch_bounds__(256) k_zdist(const Tile *__restrict__ tiles, const uint32_t *__restrict__ in,
                                               uint32_t *__restrict__ out, uint32_t *__restrict__ tilemax) {
    const Tile t = tiles[blockIdx.y];
    const int W = t.W, H = t.H, D = t.D;
    const long long HW = (long long)H * W, npix = (long long)D * HW;
    uint32_t mymax = 0;
    for (long long i0 = (long long)blockIdx.x * blockDim.x; i0 < npix; i0 += (long long)gridDim.x * blockDim.x) {
        long long i = i0 + threadIdx.x;
        if (i < npix) {
            int x, y, z;
            unravel3(i, W, H, x, y, z);
            const uint32_t *ip = in + t.base + i;
            uint32_t best = ip[0];
            for (int dz = 1;; dz++) {
                uint32_t dd = (uint32_t)dz * dz;
                if (dd >= best) break;
                bool any = false;
                if (z - dz >= 0) {
                    any = true;
                    best = min(best, ip[-(long long)dz * HW] + dd);
                }
                if (z + dz < D) {
                    any = true;
                    best = min(best, ip[(long long)dz * HW] + dd);
                }
                if (!any) break;
            }
            if (best >= DBIG) best = (uint32_t)(z + 1) * (z + 1) + (uint32_t)y * y + (uint32_t)x * x;
            out[t.base + i] = best;
            mymax = max(mymax, best);
        }
    }
    mymax = warp_max_u32(mymax);
    if ((threadIdx.x & 31) == 0 && mymax) atomicMax(&tilemax[blockIdx.y], mymax);
}

// ------------------------------------------------------------------ maximum filter (scipy, mode='reflect')
// 1-D pass along `axis` (0 = z, 1 = y, 2 = x): window [c - size/2, c - size/2 + size - 1].
// last = 1: compare with d2 and emit the seed parent array instead of the filtered value.
__global__ void __launch_bounds__(256) k_maxfilt(const Tile *__restrict__ tiles, const uint32_t *__restrict__ in,
                                                 uint32_t *__restrict__ out, int axis, int size, int last,
                                                 const uint32_t *__restrict__ d2, const uint8_t *__restrict__ msk,
                                                 uint32_t *__restrict__ par, uint32_t *__restrict__ sbits) {
    const Tile t = tiles[blockIdx.y];
    const int W = t.W, H = t.H, D = t.D;
    const long long HW = (long long)H * W, npix = (long long)D * HW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
        int x, y, z;
        unravel3(i, W, H, x, y, z);
        int c, L;
        long long st;
        if (axis == 2) {
            c = x, L = W, st = 1;
        } else if (axis == 1) {
            c = y, L = H, st = W;
        } else {
            c = z, L = D, st = HW;
        }
        const uint32_t *ip = in + t.base + i;
        int lo = c - size / 2;
        uint32_t m = 0;
        for (int k = 0; k < size; k++) {
            int j = lo + k;
            while (j < 0 || j >= L) {
                if (j < 0) j = -j - 1;
                if (j >= L) j = 2 * L - j - 1;
            }
            m = max(m, ip[(long long)(j - c) * st]);
        }
        if (!last) {
            out[t.base + i] = m;
        } else {
            bool seed = (m == d2[t.base + i]) && msk[t.base + i];
            par[t.base + i] = seed ? (uint32_t)i : NONE32;
            if (seed) atomicOr(&sbits[(t.base + i) >> 5], 1u << ((t.base + i) & 31));
        }
    }
}

// ---- shared-memory variants for 2-D slices (the default fragments_in_xy mode) ------------------------------
// Column pass over a strip of CS_W columns: the strip's row distances sit in shared memory, so the pruned
// search costs shared-memory reads instead of L1/L2 round trips.  Same arithmetic as k_coldist.
static constexpr int CS_W = 32;
__global__ void __launch_bounds__(256) k_coldist_strip(const Tile *__restrict__ tiles, const uint16_t *__restrict__ g,
                                                       uint32_t *__restrict__ out, uint32_t *__restrict__ tilemax) {
    extern __shared__ uint32_t cs_g2[];   // [H][CS_W] squared row distances (DBIG: no background in the row)
    const Tile t = tiles[blockIdx.y];
    const int W = t.W, H = t.H;
    const int nstrips = (W + CS_W - 1) / CS_W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool final2d = (t.ndim == 2);
    uint32_t mymax = 0;
    // blockIdx.x enumerates (slice z, strip)
    for (int job = blockIdx.x; job < t.D * nstrips; job += gridDim.x) {
        const int z = job / nstrips, x0 = (job - z * nstrips) * CS_W;
        const int x = x0 + lane;
        const long long sbase = t.base + (long long)z * H * W;
        __syncthreads();
        for (int y = warp; y < H; y += 8) {
            uint32_t gv = x < W ? g[sbase + (long long)y * W + x] : GINF;
            cs_g2[y * CS_W + lane] = gv == GINF ? DBIG : gv * gv;
        }
        __syncthreads();
        if (x < W)
            for (int ya = 4 * warp; ya < H; ya += 32) {
                // four vertically adjacent pixels share every row they look at: the row k above the group lies at distance
                // k + j from its j-th pixel, the row k below at distance k + 3 - j
                const uint32_t *col = cs_g2 + lane;
                uint32_t g0[4], b[4];
#pragma unroll
                for (int j = 0; j < 4; j++) g0[j] = ya + j < H ? col[(ya + j) * CS_W] : DBIG;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    b[j] = g0[j];
#pragma unroll
                    for (int i2 = 0; i2 < 4; i2++)
                        if (i2 != j) b[j] = min(b[j], g0[i2] + (uint32_t)((i2 - j) * (i2 - j)));
                }
                for (int k = 1;; k++) {
                    const uint32_t kk = (uint32_t)k * k;
                    const int ru = ya - k, rd = ya + 3 + k;
                    // rows of the group beyond the tile never win: their pixels are not stored
                    uint32_t bm = b[0];
#pragma unroll
                    for (int j = 1; j < 4; j++)
                        if (ya + j < H) bm = max(bm, b[j]);
                    if (kk >= bm || (ru < 0 && rd >= H)) break;
                    if (ru >= 0) {
                        const uint32_t g2 = col[ru * CS_W];
#pragma unroll
                        for (int j = 0; j < 4; j++) b[j] = min(b[j], g2 + kk + (uint32_t)(2 * k * j + j * j));
                    }
                    if (rd < H) {
                        const uint32_t g2 = col[rd * CS_W];
#pragma unroll
                        for (int j = 0; j < 4; j++) b[j] = min(b[j], g2 + kk + (uint32_t)(2 * k * (3 - j) + (3 - j) * (3 - j)));
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (ya + j < H) {
                        uint32_t v = b[j];
                        if (v >= DBIG) v = final2d ? (uint32_t)(ya + j + 1) * (ya + j + 1) + (uint32_t)x * x : DBIG;
                        out[sbase + (long long)(ya + j) * W + x] = v;
                        if (final2d) mymax = max(mymax, v);
                    }
            }
    }
    if (final2d) {
        mymax = warp_max_u32(mymax);
        if (lane == 0 && mymax) atomicMax(&tilemax[blockIdx.y], mymax);
    }
}

// x and y passes of the maximum filter fused through a shared-memory patch with reflected halo.
// final2d = 1: emit the seed parent array / flags (2-D tiles); 0: write the xy-filtered value (3-D tiles,
// the z pass follows with k_maxfilt).
static constexpr int MF_TW = 64, MF_TH = 32;
__device__ __forceinline__ int reflect_idx(int j, int L) {
    while (j < 0 || j >= L) {
        if (j < 0) j = -j - 1;
        if (j >= L) j = 2 * L - j - 1;
    }
    return j;
}
__global__ void __launch_bounds__(256) k_maxfilt_xy(const Tile *__restrict__ tiles, const uint32_t *__restrict__ d2, int size,
                                                    int final2d, const uint8_t *__restrict__ msk, uint32_t *__restrict__ out,
                                                    uint32_t *__restrict__ sbits) {
    extern __shared__ uint32_t mf_s[];
    const Tile t = tiles[blockIdx.y];
    const int W = t.W, H = t.H;
    const int lo = size / 2;
    const int PW = MF_TW + size - 1, PH = MF_TH + size - 1;
    uint32_t *A = mf_s;              // [PH][PW] input patch
    uint32_t *B = mf_s + PH * PW;    // [PH][MF_TW] row maxima
    const int ntx = (W + MF_TW - 1) / MF_TW, nty = (H + MF_TH - 1) / MF_TH;
    const int njobs = t.D * ntx * nty;
    for (int job = blockIdx.x; job < njobs; job += gridDim.x) {
        const int z = job / (ntx * nty), r = job - z * ntx * nty;
        const int ty = r / ntx, tx = r - ty * ntx;
        const int x0 = tx * MF_TW, y0 = ty * MF_TH;
        const long long sbase = t.base + (long long)z * H * W;
        __syncthreads();
        for (int i = threadIdx.x; i < PH * PW; i += 256) {
            int py = i / PW, px = i - py * PW;
            int yy = reflect_idx(y0 - lo + py, H), xx = reflect_idx(x0 - lo + px, W);
            A[i] = d2[sbase + (long long)yy * W + xx];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < PH * MF_TW; i += 256) {
            int py = i / MF_TW, px = i - py * MF_TW;
            const uint32_t *a = A + py * PW + px;
            uint32_t m = 0;
            for (int k = 0; k < size; k++) m = max(m, a[k]);
            B[i] = m;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < MF_TH * MF_TW; i += 256) {
            int py = i / MF_TW, px = i - py * MF_TW;
            int y = y0 + py, x = x0 + px;
            if (y < H && x < W) {
                const uint32_t *b = B + py * MF_TW + px;
                uint32_t m = 0;
                for (int k = 0; k < size; k++) m = max(m, b[k * MF_TW]);
                const long long p = sbase + (long long)y * W + x;
                if (!final2d) {
                    out[p] = m;
                } else {
                    bool seed = (m == A[(py + lo) * PW + px + lo]) && msk[p];
                    out[p] = seed ? (uint32_t)((long long)z * H * W + (long long)y * W + x) : NONE32;
                    if (seed) atomicOr(&sbits[p >> 5], 1u << (p & 31));
                }
            }
        }
    }
}

// The same for a compile-time window (the default min_seed_distance = 10): every thread produces 4 neighbouring
// outputs of a pass from SIZE + 3 shared-memory reads, which takes the kernel off the shared-memory bandwidth limit.
template <int SIZE>
__global__ void __launch_bounds__(256) k_maxfilt_xy_t(const Tile *__restrict__ tiles, const uint32_t *__restrict__ d2, int final2d,
                                                      const uint8_t *__restrict__ msk, uint32_t *__restrict__ out,
                                                      uint32_t *__restrict__ sbits) {
    static_assert(SIZE >= 4 && MF_TW % 4 == 0 && MF_TH % 4 == 0, "window / tile shape");
    constexpr int LO = SIZE / 2, PW = MF_TW + SIZE - 1, PH = MF_TH + SIZE - 1;
    constexpr int AW = (PH * PW + 3) & ~3;
    __shared__ __align__(16) uint32_t A[AW];            // [PH][PW] input patch
    __shared__ __align__(16) uint32_t B[PH * MF_TW];    // [PH][MF_TW] row maxima
    const Tile t = tiles[blockIdx.y];
    const int W = t.W, H = t.H;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ntx = (W + MF_TW - 1) / MF_TW, nty = (H + MF_TH - 1) / MF_TH;
    const int njobs = t.D * ntx * nty;
    for (int job = blockIdx.x; job < njobs; job += gridDim.x) {
        const int z = job / (ntx * nty), r = job - z * ntx * nty;
        const int ty = r / ntx, tx = r - ty * ntx;
        const int x0 = tx * MF_TW, y0 = ty * MF_TH;
        const long long sbase = t.base + (long long)z * H * W;
        __syncthreads();
        for (int py = warp; py < PH; py += 8) {
            const long long rowb = sbase + (long long)reflect_idx(y0 - LO + py, H) * W;
            for (int px = lane; px < PW; px += 32) A[py * PW + px] = d2[rowb + reflect_idx(x0 - LO + px, W)];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < PH * (MF_TW / 4); i += 256) {
            const int py = i / (MF_TW / 4), q = i % (MF_TW / 4);
            const uint32_t *a = A + py * PW + 4 * q;
            uint32_t v[SIZE + 3];
#pragma unroll
            for (int k = 0; k < SIZE + 3; k++) v[k] = a[k];
            uint32_t c = v[3];
#pragma unroll
            for (int k = 4; k < SIZE; k++) c = max(c, v[k]);
            uint4 o;
            o.x = max(max(c, v[0]), max(v[1], v[2]));
            o.y = max(max(c, v[1]), max(v[2], v[SIZE]));
            o.z = max(max(c, v[2]), max(v[SIZE], v[SIZE + 1]));
            o.w = max(max(c, v[SIZE]), max(v[SIZE + 1], v[SIZE + 2]));
            *reinterpret_cast<uint4 *>(B + py * MF_TW + 4 * q) = o;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < (MF_TH / 4) * MF_TW; i += 256) {
            const int pq = i / MF_TW, px = i % MF_TW;
            const uint32_t *b = B + (4 * pq) * MF_TW + px;
            uint32_t v[SIZE + 3];
#pragma unroll
            for (int k = 0; k < SIZE + 3; k++) v[k] = b[k * MF_TW];
            uint32_t c = v[3];
#pragma unroll
            for (int k = 4; k < SIZE; k++) c = max(c, v[k]);
            uint32_t o[4];
            o[0] = max(max(c, v[0]), max(v[1], v[2]));
            o[1] = max(max(c, v[1]), max(v[2], v[SIZE]));
            o[2] = max(max(c, v[2]), max(v[SIZE], v[SIZE + 1]));
            o[3] = max(max(c, v[SIZE]), max(v[SIZE + 1], v[SIZE + 2]));
            const int x = x0 + px;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int py = 4 * pq + j, y = y0 + py;
                if (y < H && x < W) {
                    const long long p = sbase + (long long)y * W + x;
                    if (!final2d) {
                        out[p] = o[j];
                    } else {
                        bool seed = (o[j] == A[(py + LO) * PW + px + LO]) && msk[p];
                        out[p] = seed ? (uint32_t)((long long)z * H * W + (long long)y * W + x) : NONE32;
                        if (seed) atomicOr(&sbits[p >> 5], 1u << (p & 31));
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------ seed connected components (conn-1)
__global__ void __launch_bounds__(256) k_seed_union(const Tile *__restrict__ tiles, uint32_t *__restrict__ par) {
    const Tile t = tiles[blockIdx.y];
    const int W = t.W, H = t.H, D = t.D;
    const long long HW = (long long)H * W, npix = (long long)D * HW;
    uint32_t *pp = par + t.base;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
        if (__ldcg(&pp[i]) == NONE32) continue;
        int x, y, z;
        unravel3f((uint32_t)i, W, H, t.fW, t.fH, x, y, z);
        if (x > 0 && __ldcg(&pp[i - 1]) != NONE32) uf_union(pp, (uint32_t)i, (uint32_t)(i - 1));
        if (y > 0 && __ldcg(&pp[i - W]) != NONE32) uf_union(pp, (uint32_t)i, (uint32_t)(i - W));
        if (z > 0 && __ldcg(&pp[i - HW]) != NONE32) uf_union(pp, (uint32_t)i, (uint32_t)(i - HW));
    }
}

// lab = root+1 for seeds, UNLAB in mask, 0 outside; histogram of d2 over the mask
__global__ void __launch_bounds__(256) k_seed_label_hist(const Tile *__restrict__ tiles, const uint32_t *__restrict__ par,
                                                         const uint8_t *__restrict__ msk, const uint32_t *__restrict__ d2,
                                                         uint32_t *__restrict__ lab, const uint32_t *__restrict__ hbase,
                                                         uint32_t *__restrict__ hist, const uint32_t *__restrict__ sbits,
                                                         const uint32_t *__restrict__ swscan, int compact_labels) {
    const Tile t = tiles[blockIdx.y];
    const long long npix = (long long)t.D * t.H * t.W;
    const uint32_t *pp = par + t.base;
    const uint32_t hb = hbase[blockIdx.y];
    const uint32_t s0 = compact_labels ? bit_rank(sbits, swscan, (uint32_t)t.base) : 0;
    // most mask pixels carry a small d2: privatise the low part of the histogram per CTA
    constexpr int SH = 2048;
    __shared__ uint32_t sh[SH];
    for (int j = threadIdx.x; j < SH; j += blockDim.x) sh[j] = 0;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
        uint32_t l = 0;
        if (msk[t.base + i]) {
            l = UNLAB;
            if (pp[i] != NONE32) {
                uint32_t root = uf_find(pp, (uint32_t)i);
                // label = root pixel + 1, or (flood v2) the root's rank among the tile's seed pixels + 1
                l = compact_labels ? bit_rank(sbits, swscan, (uint32_t)(t.base + root)) - s0 + 1 : root + 1;
            }
            uint32_t d = d2[t.base + i];
            if (d < SH)
                atomicAdd(&sh[d], 1u);
            else
                atomicAdd(&hist[hb + d], 1u);
        }
        // compact_labels == 2: the label plane is rewritten as a whole after the flood (k_scatter_labels_tile), only the
        // seeds' labels are read before that
        if (compact_labels != 2 || (l != 0 && l != UNLAB)) lab[t.base + i] = l;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < SH; j += blockDim.x)
        if (sh[j]) atomicAdd(&hist[hb + j], sh[j]);
}

__global__ void k_tile_hsize(const uint32_t *__restrict__ tilemax, uint32_t *__restrict__ hsize, int ntiles) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ntiles) hsize[i] = tilemax[i] + 1;
}

__global__ void k_nzflag(const uint32_t *__restrict__ hist, uint8_t *__restrict__ nz, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) nz[i] = hist[i] ? 1 : 0;
}

// level tables: for every non-empty histogram entry, its dense rank gets the FIFO segment start
__global__ void k_levels(const uint32_t *__restrict__ hist, const uint32_t *__restrict__ lrank,
                         const uint32_t *__restrict__ qoff, uint32_t *__restrict__ lvl_qstart, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && hist[i]) lvl_qstart[lrank[i]] = qoff[i];
}

__global__ void k_tile_ranges(const Tile *__restrict__ tiles, int ntiles, const uint32_t *__restrict__ hbase,
                              const uint32_t *__restrict__ lrank, const uint32_t *__restrict__ nlevels,
                              const uint32_t *__restrict__ sbits, const uint32_t *__restrict__ swscan,
                              const uint32_t *__restrict__ nseeds, uint32_t *__restrict__ tile_lvl, uint32_t *__restrict__ tile_seed,
                              const uint32_t *__restrict__ qoff, const uint32_t *__restrict__ nmask,
                              uint32_t *__restrict__ tile_q) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ntiles) {
        tile_lvl[i] = lrank[hbase[i]];
        tile_seed[i] = bit_rank(sbits, swscan, (uint32_t)tiles[i].base);
        tile_q[i] = qoff[hbase[i]];
    } else if (i == ntiles) {
        tile_lvl[i] = *nlevels;
        tile_seed[i] = *nseeds;
        tile_q[i] = *nmask;
    }
}

// largest number of priority levels in one tile (sizes the shared-memory level tables of flood v2)
__global__ void k_tile_nlev_max(const uint32_t *__restrict__ tile_lvl, int ntiles, uint32_t *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ntiles) atomicMax(out, tile_lvl[i + 1] - tile_lvl[i]);
}

// largest number of seed pixels in one tile (decides whether 15-bit labels suffice for flood v2)
__global__ void k_tile_seed_max(const Tile *__restrict__ tiles, int ntiles, const uint32_t *__restrict__ sbits,
                                const uint32_t *__restrict__ swscan,
                                const uint32_t *__restrict__ nseeds, const uint32_t *__restrict__ tilemax,
                                uint32_t *__restrict__ out_seedmax, uint32_t *__restrict__ out_d2max) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ntiles) return;
    uint32_t b = bit_rank(sbits, swscan, (uint32_t)tiles[i].base);
    uint32_t e = i + 1 < ntiles ? bit_rank(sbits, swscan, (uint32_t)tiles[i + 1].base) : *nseeds;
    atomicMax(out_seedmax, e - b);
    atomicMax(out_d2max, tilemax[i]);
}

// lv = dense level rank of every mask pixel; seed list compaction (tile-local pixel indices)
__global__ void __launch_bounds__(256) k_pixel_levels(const Tile *__restrict__ tiles, const uint8_t *__restrict__ msk,
                                                      const uint32_t *__restrict__ d2, const uint32_t *__restrict__ hbase,
                                                      const uint32_t *__restrict__ lrank, uint32_t *__restrict__ lv,
                                                      const uint32_t *__restrict__ sbits, const uint32_t *__restrict__ swscan,
                                                      uint32_t *__restrict__ seedlist, uint16_t *__restrict__ lv16,
                                                      uint32_t *__restrict__ availw) {
    const Tile t = tiles[blockIdx.y];
    const long long npix = (long long)t.D * t.H * t.W;
    const uint32_t hb = hbase[blockIdx.y];
    const uint32_t l0 = lrank[hb];
    for (long long i0 = (long long)blockIdx.x * blockDim.x; i0 < npix; i0 += (long long)gridDim.x * blockDim.x) {
        long long i = i0 + threadIdx.x;
        long long p = t.base + i;
        bool av = false;
        if (i < npix) {
            bool m = msk[p] != 0, sd = bit_test(sbits, (uint32_t)p);
            if (m) {
                uint32_t r = lrank[hb + d2[p]];
                if (lv16)
                    lv16[p] = (uint16_t)(r - l0);   // tile-relative level (flood v2)
                else
                    lv[p] = r;
            }
            if (sd) seedlist[bit_rank(sbits, swscan, (uint32_t)p)] = (uint32_t)i;
            av = m && !sd;
        }
        if (availw) {
            // tile bases are 32-aligned: one ballot = one word of the tile's "in mask, not labelled" bitmap
            unsigned w = __ballot_sync(FULL, av);
            if ((threadIdx.x & 31) == 0 && i < npix) availw[p >> 5] = w;
        }
    }
}

// ------------------------------------------------------------------ the flood
// Order-preserving append of up to NS candidates per lane (order: lane-major, slot-minor) to the
// FIFO of their level.  `cur`/`tailc`: the current level's tail lives in a register.
// ST: the tails of the tile's levels live in shared memory (stail[level - lo]) together with a bitmap of the levels that
// hold entries (snz); otherwise they are read / written through L2 (lvl_tail).
template <int NS, bool ST>
__device__ __forceinline__ void warp_append(const bool (&valid_in)[NS], const uint32_t (&lvl)[NS], const uint32_t (&pix)[NS],
                                            uint32_t *__restrict__ queue, const uint32_t *__restrict__ lvl_qstart,
                                            uint32_t *__restrict__ lvl_tail, uint32_t cur, uint32_t &tailc, int lane,
                                            uint32_t *stail = nullptr, uint32_t *snz = nullptr, uint32_t lo = 0) {
    bool valid[NS];
    uint32_t tl[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) {
        valid[s] = valid_in[s];
        tl[s] = 0;
        if (valid[s] && lvl[s] != cur) tl[s] = ST ? stail[lvl[s] - lo] : __ldcg(&lvl_tail[lvl[s]]);
    }
    for (;;) {
        uint32_t mylo = NONE32;
#pragma unroll
        for (int s = 0; s < NS; s++)
            if (valid[s]) mylo = min(mylo, lvl[s]);
        uint32_t L = __reduce_min_sync(FULL, mylo);
        if (L == NONE32) break;
        int c = 0;
        uint32_t mytail = 0;
#pragma unroll
        for (int s = NS - 1; s >= 0; s--)
            if (valid[s] && lvl[s] == L) {
                c++;
                mytail = tl[s];
            }
        unsigned has = __ballot_sync(FULL, c > 0);
        int src = __ffs(has) - 1;
        uint32_t base = __shfl_sync(FULL, mytail, src);
        if (L == cur) base = tailc;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += v;
        }
        int total = __shfl_sync(FULL, incl, 31);
        uint32_t pos = __ldg(&lvl_qstart[L]) + base + (uint32_t)(incl - c);
#pragma unroll
        for (int s = 0; s < NS; s++)
            if (valid[s] && lvl[s] == L) {
                __stcg(&queue[pos++], pix[s]);
                valid[s] = false;
            }
        if (L == cur)
            tailc = base + total;
        else if (lane == 0) {
            if (ST) {
                stail[L - lo] = base + (uint32_t)total;
                snz[(L - lo) >> 5] |= 1u << ((L - lo) & 31);
            } else
                __stcg(&lvl_tail[L], base + (uint32_t)total);
        }
        if (ST) __syncwarp();
    }
}

// One warp per tile.  Level ranks: higher rank = larger d2 = smaller skimage value = pops first.
__global__ void __launch_bounds__(64) k_flood(const Tile *__restrict__ tiles, int ntiles, uint32_t *__restrict__ lab_all,
                                              const uint32_t *__restrict__ lv_all, uint32_t *__restrict__ queue,
                                              const uint32_t *__restrict__ lvl_qstart, uint32_t *__restrict__ lvl_head,
                                              uint32_t *__restrict__ lvl_tail, const uint32_t *__restrict__ tile_lvl,
                                              const uint32_t *__restrict__ seedlist, const uint32_t *__restrict__ tile_seed,
                                              uint32_t *__restrict__ stats) {
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= ntiles) return;
    const Tile t = tiles[wid];
    uint32_t *lab = lab_all + t.base;
    const uint32_t *lv = lv_all + t.base;
    const int W = t.W, H = t.H, D = t.D;
    const int HW = H * W;
    const uint32_t lo = tile_lvl[wid], hi = tile_lvl[wid + 1];
    if (lo == hi) return;
    uint32_t steps = 0, intr = 0;

    // ---- seeds, ascending raveled index (oracle seed_tie = "index")
    uint32_t cur = lo, tailc = 0, headc = 0;
    {
        const uint32_t sb = tile_seed[wid], se = tile_seed[wid + 1];
        if (sb == se) return;
        uint32_t maxr = lo;
        uint32_t dummy_tail = 0;
        for (uint32_t s0 = sb; s0 < se; s0 += 32) {
            bool v[1];
            uint32_t l[1], px[1];
            v[0] = s0 + lane < se;
            px[0] = v[0] ? seedlist[s0 + lane] : 0;
            l[0] = v[0] ? lv[px[0]] : 0;
            if (v[0]) maxr = max(maxr, l[0]);
            warp_append<1, false>(v, l, px, queue, lvl_qstart, lvl_tail, NONE32, dummy_tail, lane);
            __syncwarp();
        }
        cur = __reduce_max_sync(FULL, maxr);
        headc = 0;
        tailc = __ldcg(&lvl_tail[cur]);
    }
    uint32_t qs = lvl_qstart[cur];

    for (;;) {
        if (headc == tailc) {
            // level exhausted: write back and descend to the next non-empty level
            if (lane == 0) {
                __stcg(&lvl_head[cur], headc);
                __stcg(&lvl_tail[cur], tailc);
            }
            __syncwarp();
            bool found = false;
            long long r0 = (long long)cur - 1;
            while (r0 >= (long long)lo) {
                long long r = r0 - lane;
                bool ne = false;
                if (r >= (long long)lo) ne = __ldcg(&lvl_tail[r]) != __ldcg(&lvl_head[r]);
                unsigned b = __ballot_sync(FULL, ne);
                if (b) {
                    cur = (uint32_t)(r0 - (__ffs(b) - 1));
                    found = true;
                    break;
                }
                r0 -= 32;
            }
            if (!found) break;
            headc = __ldcg(&lvl_head[cur]);
            tailc = __ldcg(&lvl_tail[cur]);
            qs = lvl_qstart[cur];
            continue;
        }
        steps++;
        const uint32_t k = min(32u, tailc - headc);
        const bool act = (uint32_t)lane < k;
        uint32_t p = 0, mylab = 0;
        if (act) {
            p = __ldcg(&queue[qs + headc + lane]);
            mylab = __ldcg(&lab[p]);
        }
        int z = p / HW;
        int rem = p - z * HW;
        int y = rem / W;
        int x = rem - y * W;
        // neighbour order of skimage (connectivity 1): -z, -y, -x, +x, +y, +z
        uint32_t nb[6];
        bool cand[6];
        nb[0] = (act && z > 0) ? p - HW : NONE32;
        nb[1] = (act && y > 0) ? p - W : NONE32;
        nb[2] = (act && x > 0) ? p - 1 : NONE32;
        nb[3] = (act && x + 1 < W) ? p + 1 : NONE32;
        nb[4] = (act && y + 1 < H) ? p + W : NONE32;
        nb[5] = (act && z + 1 < D) ? p + HW : NONE32;
        const uint32_t keybase = CLAIM | ((uint32_t)lane << 3);
#pragma unroll
        for (int s = 0; s < 6; s++) {
            cand[s] = false;
            if (nb[s] != NONE32) {
                uint32_t old = atomicMin(&lab[nb[s]], keybase | s);
                cand[s] = old >= CLAIM;
            }
        }
        __syncwarp();
        uint32_t l[6];
        bool up = false;
#pragma unroll
        for (int s = 0; s < 6; s++) {
            l[s] = 0;
            if (cand[s]) {
                cand[s] = __ldcg(&lab[nb[s]]) == (keybase | s);
                if (cand[s]) {
                    l[s] = lv[nb[s]];
                    up |= l[s] > cur;
                }
            }
        }
        const unsigned ball = __ballot_sync(FULL, up);
        const int rstar = ball ? __ffs(ball) - 1 : 31;
        if (lane > rstar) {
#pragma unroll
            for (int s = 0; s < 6; s++)
                if (cand[s]) {
                    __stcg(&lab[nb[s]], UNLAB);
                    cand[s] = false;
                }
        } else {
#pragma unroll
            for (int s = 0; s < 6; s++)
                if (cand[s]) __stcg(&lab[nb[s]], mylab);
        }
        headc += min(k, (uint32_t)rstar + 1u);
        warp_append<6, false>(cand, l, nb, queue, lvl_qstart, lvl_tail, cur, tailc, lane);
        __syncwarp();
        if (ball) {
            intr++;
            uint32_t mx = 0;
#pragma unroll
            for (int s = 0; s < 6; s++)
                if (cand[s]) mx = max(mx, l[s]);
            mx = __reduce_max_sync(FULL, mx);
            if (lane == 0) {
                __stcg(&lvl_head[cur], headc);
                __stcg(&lvl_tail[cur], tailc);
            }
            __syncwarp();
            cur = mx;
            headc = __ldcg(&lvl_head[cur]);
            tailc = __ldcg(&lvl_tail[cur]);
            qs = lvl_qstart[cur];
        }
    }
    if (lane == 0 && stats) {
        atomicAdd(&stats[0], steps);
        atomicAdd(&stats[1], intr);
        atomicMax(&stats[2], steps);
    }
}

// ------------------------------------------------------------------ faithful flood (bs_set_flood_version(6))
// skimage's watershed loop replayed literally, one thread per tile: a binary heap ordered by (value, age) -- here (level
// descending, age ascending) -- with skimage's own layout rules (parent = (child + 1) / 2 - 1; pop = swap the last element
// to the root, sift down preferring the left child unless the right one is strictly smaller than the smaller so far), all
// seeds pushed with age 0 in ascending raveled index.  Equal-valued seeds therefore leave the heap in the order its
// layout history dictates, which is what the reference does and what the default floods replace by the index rule
// (declared deviation D1).  Sequential and latency bound (~10^2 slower than the default flood): for parity runs.
struct HeapRef {
    uint32_t *lv, *age, *pix;
};
__device__ __forceinline__ bool heap_smaller(const HeapRef &h, uint32_t a, uint32_t b) {
    const uint32_t la = h.lv[a], lb = h.lv[b];
    if (la != lb) return la > lb;          // larger level = smaller skimage value
    return h.age[a] < h.age[b];
}
__device__ __forceinline__ void heap_swap(const HeapRef &h, uint32_t a, uint32_t b) {
    const uint32_t l = h.lv[a], g = h.age[a], p = h.pix[a];
    h.lv[a] = h.lv[b], h.age[a] = h.age[b], h.pix[a] = h.pix[b];
    h.lv[b] = l, h.age[b] = g, h.pix[b] = p;
}
__global__ void __launch_bounds__(32) k_flood_heap(const Tile *__restrict__ tiles, int ntiles, uint32_t *__restrict__ lab_all,
                                                   const uint32_t *__restrict__ lv_all, const uint32_t *__restrict__ seedlist,
                                                   const uint32_t *__restrict__ tile_seed, uint32_t *__restrict__ hlv,
                                                   uint32_t *__restrict__ hage, uint32_t *__restrict__ hpix) {
    const int wid = blockIdx.x * blockDim.x + threadIdx.x;
    if (wid >= ntiles) return;
    const Tile t = tiles[wid];
    uint32_t *lab = lab_all + t.base;
    const uint32_t *lv = lv_all + t.base;
    HeapRef h;
    h.lv = hlv + t.base, h.age = hage + t.base, h.pix = hpix + t.base;
    const int W = t.W, H = t.H, D = t.D;
    const uint32_t HW = (uint32_t)H * W;
    uint32_t n = 0;
    auto push = [&](uint32_t level, uint32_t age, uint32_t p) {
        uint32_t child = n++;
        h.lv[child] = level, h.age[child] = age, h.pix[child] = p;
        while (child > 0) {
            const uint32_t parent = (child + 1) / 2 - 1;
            if (!heap_smaller(h, child, parent)) break;
            heap_swap(h, child, parent);
            child = parent;
        }
    };
    for (uint32_t s0 = tile_seed[wid]; s0 < tile_seed[wid + 1]; s0++) {
        const uint32_t p = seedlist[s0];
        push(lv[p], 0u, p);
    }
    uint32_t age = 1;
    while (n > 0) {
        const uint32_t p = h.pix[0];
        n--;
        if (n > 0) {
            heap_swap(h, 0, n);
            uint32_t i = 0, smallest = 0;
            for (;;) {
                const uint32_t l = 2 * i + 1, r = 2 * i + 2;
                if (l >= n) break;
                if (heap_smaller(h, l, i)) smallest = l;
                if (r < n && heap_smaller(h, r, smallest)) smallest = r;
                if (smallest == i) break;
                heap_swap(h, i, smallest);
                i = smallest;
            }
        }
        const uint32_t mylab = lab[p];
        const int z = (int)(p / HW);
        const uint32_t rem = p - (uint32_t)z * HW;
        const int y = (int)(rem / (uint32_t)W), x = (int)(rem - (uint32_t)y * W);
        // neighbour order of skimage (connectivity 1): -z, -y, -x, +x, +y, +z
        uint32_t nb[6];
        nb[0] = z > 0 ? p - HW : NONE32;
        nb[1] = y > 0 ? p - W : NONE32;
        nb[2] = x > 0 ? p - 1 : NONE32;
        nb[3] = x + 1 < W ? p + 1 : NONE32;
        nb[4] = y + 1 < H ? p + W : NONE32;
        nb[5] = z + 1 < D ? p + HW : NONE32;
        for (int k = 0; k < 6; k++) {
            const uint32_t q = nb[k];
            if (q == NONE32 || lab[q] != UNLAB) continue;    // outside the mask or labelled already
            age++;
            lab[q] = mylab;                                  // labelled at push time
            push(lv[q], age, q);
        }
    }
}

// ------------------------------------------------------------------ flood v3 (any tile; one CTA per tile)
// The step semantics of k_flood with F3_NT queue entries per step instead of 32: tiles that flood v2 cannot take (3-D
// read ROIs, slices beyond 2^17 pixels) hold millions of pixels in a few thousand levels, so a level's FIFO feeds a whole
// CTA.  Claims resolve by the lowest (thread, slot) key -- exactly what the lowest (lane, slot) key does in one warp --
// in a shared-memory table keyed by pixel, so nothing is written to global memory before the step's cut is known; the
// step is cut after the first entry (in FIFO order) that queued a pixel above the current level; the ordered append
// goes through a per-step table level -> (queue position, entries per warp).  Heads, tails and an occupancy bitmap of
// the tile's levels live in shared memory: per level and step the dependent global accesses are the entry, the labels of
// its neighbours and their levels.
static constexpr int F3_NT = 512, F3_NW = F3_NT / 32, F3_SLOTS = 1024, F3_CLAIMS = 8192;
static constexpr uint32_t F3_PIXMASK = (1u << 29) - 1u;
struct F3Shared {
    uint32_t hk[F3_SLOTS];             // level of the slot (NONE32: free)
    uint32_t hbase[F3_SLOTS];          // queue position where this step's entries of the level start
    uint16_t hcnt[F3_SLOTS][F3_NW];    // entries per warp, then their exclusive prefix over the warps
    uint32_t ck[F3_CLAIMS];            // claim table: pixel ...
    uint32_t cm[F3_CLAIMS];            // ... -> lowest (thread * 8 + slot) that wants it
    uint32_t rstar, mx, tailc, next;
};
static size_t flood3_smem(int levcap) { return sizeof(F3Shared) + 4 * (size_t)(2 * levcap + (levcap + 31) / 32 + 1); }

// ordered (thread-major, slot-minor) append of the block's candidates to the FIFOs of their levels.  All threads call it.
// Every warp ranks its candidates level by level (a step touches few distinct levels per warp) and leaves one count per
// (level slot, warp); a prefix over the warps then places the warps' runs behind the level's tail.
// On return `tailc` is the tail of level `cur` and `mx_out` the value of S.mx.
template <int NS>
__device__ __forceinline__ void block_append(F3Shared &S, const bool (&valid)[NS], const uint32_t (&lvl)[NS], const uint32_t (&pix)[NS],
                                             uint32_t *__restrict__ queue, const uint32_t *__restrict__ lvl_qstart,
                                             uint32_t *ltail, uint32_t *lnz, uint32_t lo, uint32_t cur, uint32_t &tailc,
                                             uint32_t &mx_out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    bool rem[NS];
    uint32_t rank[NS], slot[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) rem[s] = valid[s], rank[s] = 0, slot[s] = 0;
    for (;;) {
        uint32_t mylo = NONE32;
#pragma unroll
        for (int s = 0; s < NS; s++)
            if (rem[s]) mylo = min(mylo, lvl[s]);
        const uint32_t L = __reduce_min_sync(FULL, mylo);
        if (L == NONE32) break;
        int c = 0;
#pragma unroll
        for (int s = 0; s < NS; s++)
            if (rem[s] && lvl[s] == L) c++;
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(FULL, incl, 31);
        uint32_t sl = 0;
        if (lane == 0) {
            uint32_t h = (L * 2654435761u) >> 22;   // 10 bits
            for (int probes = 0;; probes++) {
                const uint32_t old = atomicCAS(&S.hk[h], NONE32, L);
                if (old == NONE32 || old == L) break;
                if (probes > F3_SLOTS) __trap();    // more distinct levels in one step than the table holds
                h = (h + 1) & (F3_SLOTS - 1);
            }
            S.hcnt[h][warp] = (uint16_t)total;
            sl = h;
        }
        sl = __shfl_sync(FULL, sl, 0);
        uint32_t r = (uint32_t)(incl - c);
#pragma unroll
        for (int s = 0; s < NS; s++)
            if (rem[s] && lvl[s] == L) {
                rank[s] = r++;
                slot[s] = sl;
                rem[s] = false;
            }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < F3_SLOTS; j += F3_NT) {
        const uint32_t L = S.hk[j];
        if (L != NONE32) {
            uint32_t run = 0;
#pragma unroll
            for (int w = 0; w < F3_NW; w++) {
                const uint32_t cw = S.hcnt[j][w];
                S.hcnt[j][w] = (uint16_t)run;
                run += cw;
            }
            const uint32_t base = L == cur ? tailc : ltail[L - lo];
            S.hbase[j] = __ldg(&lvl_qstart[L]) + base;
            if (L == cur)
                S.tailc = base + run;
            else {
                ltail[L - lo] = base + run;
                atomicOr(&lnz[(L - lo) >> 5], 1u << ((L - lo) & 31));
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < NS; s++)
        if (valid[s]) __stcg(&queue[S.hbase[slot[s]] + S.hcnt[slot[s]][warp] + rank[s]], pix[s]);
    tailc = S.tailc;
    mx_out = S.mx;
    __syncthreads();
    for (int j = threadIdx.x; j < F3_SLOTS; j += F3_NT)
        if (S.hk[j] != NONE32) {
            S.hk[j] = NONE32;
#pragma unroll
            for (int w = 0; w < F3_NW; w += 2) *(uint32_t *)&S.hcnt[j][w] = 0u;
        }
    // the callers pass a barrier before the table is used again
}

__global__ void __launch_bounds__(F3_NT) k_flood3(const Tile *__restrict__ tiles, int ntiles, uint32_t *__restrict__ lab_all,
                                                  const uint32_t *__restrict__ lv_all, uint32_t *__restrict__ queue,
                                                  const uint32_t *__restrict__ lvl_qstart, const uint32_t *__restrict__ tile_lvl,
                                                  const uint32_t *__restrict__ seedlist, const uint32_t *__restrict__ tile_seed,
                                                  int levcap, uint32_t *__restrict__ stats) {
    extern __shared__ __align__(16) unsigned char f3_raw[];
    F3Shared &S = *(F3Shared *)f3_raw;
    uint32_t *lhead = (uint32_t *)(f3_raw + sizeof(F3Shared)), *ltail = lhead + levcap, *lnz = ltail + levcap;
    const int wid = blockIdx.x, tid = threadIdx.x;
    const Tile t = tiles[wid];
    uint32_t *lab = lab_all + t.base;
    const uint32_t *lv = lv_all + t.base;
    const int W = t.W, H = t.H, D = t.D;
    const uint32_t HW = (uint32_t)H * W;
    const FastDiv fHW = make_fastdiv_dev(HW);
    const uint32_t lo = tile_lvl[wid], hi = tile_lvl[wid + 1];
    const uint32_t sb = tile_seed[wid], se = tile_seed[wid + 1];
    if (lo == hi || sb == se) return;
    for (int j = tid; j < F3_SLOTS; j += F3_NT) {
        S.hk[j] = NONE32;
#pragma unroll
        for (int w = 0; w < F3_NW; w++) S.hcnt[j][w] = 0;
    }
    for (int j = tid; j < F3_CLAIMS; j += F3_NT) S.ck[j] = NONE32, S.cm[j] = NONE32;
    for (int j = tid; j < 2 * levcap + (levcap + 31) / 32; j += F3_NT) lhead[j] = 0;
    if (tid == 0) S.mx = lo, S.tailc = 0, S.rstar = NONE32, S.next = 0;
    __syncthreads();
    uint32_t steps = 0, intr = 0;
    uint32_t cur = lo, tailc = 0, headc = 0, mx = 0;
    // ---- seeds, ascending raveled index (oracle seed_tie = "index")
    for (uint32_t s0 = sb; s0 < se; s0 += F3_NT) {
        bool v[1];
        uint32_t l[1], px[1];
        v[0] = s0 + tid < se;
        px[0] = v[0] ? seedlist[s0 + tid] : 0;
        l[0] = v[0] ? lv[px[0]] : 0;
        if (v[0]) atomicMax(&S.mx, l[0]);
        uint32_t dummy_tail = 0;
        block_append<1>(S, v, l, px, queue, lvl_qstart, ltail, lnz, lo, NONE32, dummy_tail, mx);
        __syncthreads();
    }
    cur = S.mx;
    headc = 0;
    tailc = ltail[cur - lo];
    uint32_t qs = lvl_qstart[cur];
    uint32_t pre_level = NONE32, pre_head = 0, pre_tail = 0, pre_e = 0;
    __syncthreads();

    for (;;) {
        if (headc == tailc) {
            // level exhausted: the next one is the highest occupied level below it (levels above are always empty)
            if (tid == 0) {
                lhead[cur - lo] = headc;
                ltail[cur - lo] = tailc;
                lnz[(cur - lo) >> 5] &= ~(1u << ((cur - lo) & 31));
                S.next = 0;
            }
            __syncthreads();
            const int r = (int)(cur - lo);
            bool found = false;
            for (int wtop = (r - 1) >> 5; r > 0 && wtop >= 0; wtop -= F3_NT) {
                const int wi = wtop - tid;
                uint32_t wv = wi >= 0 ? lnz[wi] : 0u;
                if (wi == ((r - 1) >> 5) && (r & 31)) wv &= (1u << (r & 31)) - 1u;
                if (wv) atomicMax(&S.next, (uint32_t)(wi * 32 + 31 - __clz(wv)) + 1u);
                __syncthreads();
                const uint32_t nx = S.next;
                __syncthreads();
                if (nx) {
                    cur = lo + nx - 1;
                    found = true;
                    break;
                }
            }
            if (!found) break;
            headc = lhead[cur - lo];
            tailc = ltail[cur - lo];
            qs = __ldg(&lvl_qstart[cur]);
            continue;
        }
        steps++;
        const uint32_t k = min((uint32_t)F3_NT, tailc - headc);
        const bool act = (uint32_t)tid < k;
        // entry = pixel | (slot it was claimed through + 1) << 29 (0: a seed); the entries that were queued when the
        // previous step started were fetched then
        uint32_t p = 0, mylab = 0, from = 0;
        if (act) {
            const uint32_t e = (pre_level == cur && pre_head == headc && headc + tid < pre_tail) ? pre_e : __ldcg(&queue[qs + headc + tid]);
            p = e & F3_PIXMASK;
            from = e >> 29;
        }
        // fetch ahead: the batch after this one, as far as it is queued already (used if this step consumes all F3_NT)
        pre_level = cur, pre_head = headc + F3_NT, pre_tail = tailc;
        pre_e = headc + F3_NT + tid < tailc ? __ldcg(&queue[qs + headc + F3_NT + tid]) : 0u;
        const int z = (int)fdiv(p, fHW);
        const uint32_t rem = p - (uint32_t)z * HW;
        const int y = (int)fdiv(rem, t.fW), x = (int)rem - y * W;
        // neighbour order of skimage (connectivity 1): -z, -y, -x, +x, +y, +z
        uint32_t nb[6], cslot[6], l[6];
        bool cand[6];
        nb[0] = (act && z > 0) ? p - HW : NONE32;
        nb[1] = (act && y > 0) ? p - W : NONE32;
        nb[2] = (act && x > 0) ? p - 1 : NONE32;
        nb[3] = (act && x + 1 < W) ? p + 1 : NONE32;
        nb[4] = (act && y + 1 < H) ? p + W : NONE32;
        nb[5] = (act && z + 1 < D) ? p + HW : NONE32;
        // the pixel this entry was claimed from is labelled: slot s was entered from its opposite, slot 5 - s
#pragma unroll
        for (int s = 0; s < 6; s++)
            if (from == (uint32_t)(6 - s)) nb[s] = NONE32;
        // one 16-byte load covers the pixel's own label and, mostly, both x neighbours (tile bases are 32-aligned)
        uint32_t lx0 = 0, lx1 = 0;
        if (act) {
            const uint4 v = __ldcg((const uint4 *)(lab + (p & ~3u)));
            const uint32_t q = p & 3u;
            mylab = q == 0 ? v.x : (q == 1 ? v.y : (q == 2 ? v.z : v.w));
            lx0 = q == 0 ? (nb[2] != NONE32 ? __ldcg(&lab[p - 1]) : 0u) : (q == 1 ? v.x : (q == 2 ? v.y : v.z));
            lx1 = q == 3 ? (nb[3] != NONE32 ? __ldcg(&lab[p + 1]) : 0u) : (q == 0 ? v.y : (q == 1 ? v.z : v.w));
        }
        cand[2] = nb[2] != NONE32 && lx0 == UNLAB;
        cand[3] = nb[3] != NONE32 && lx1 == UNLAB;
#pragma unroll
        for (int s = 0; s < 6; s++)
            if (s != 2 && s != 3) cand[s] = nb[s] != NONE32 && __ldcg(&lab[nb[s]]) == UNLAB;
#pragma unroll
        for (int s = 0; s < 6; s++) l[s] = cand[s] ? lv[nb[s]] : 0u;   // in flight while the claims resolve
        const uint32_t keybase = (uint32_t)tid << 3;
#pragma unroll
        for (int s = 0; s < 6; s++) {
            cslot[s] = 0;
            if (cand[s]) {
                uint32_t h = (nb[s] * 2654435761u) >> 19;   // 13 bits
                for (;;) {
                    const uint32_t old = atomicCAS(&S.ck[h], NONE32, nb[s]);
                    if (old == NONE32 || old == nb[s]) break;
                    h = (h + 1) & (F3_CLAIMS - 1);
                }
                cslot[s] = h;
                atomicMin(&S.cm[h], keybase | s);
            }
        }
        if (tid == 0) S.rstar = NONE32, S.mx = 0, S.tailc = tailc;
        __syncthreads();
        bool up = false;
        bool mine[6];
#pragma unroll
        for (int s = 0; s < 6; s++) {
            mine[s] = cand[s];
            if (cand[s]) {
                cand[s] = S.cm[cslot[s]] == (keybase | s);
                if (cand[s])
                    up |= l[s] > cur;
                else
                    l[s] = 0;
            }
        }
        if (up) atomicMin(&S.rstar, (uint32_t)tid);
        __syncthreads();
        const uint32_t rstar = S.rstar;   // NONE32: nothing above the current level was queued
#pragma unroll
        for (int s = 0; s < 6; s++)
            if (mine[s]) S.ck[cslot[s]] = NONE32, S.cm[cslot[s]] = NONE32;
        if ((uint32_t)tid > rstar) {
#pragma unroll
            for (int s = 0; s < 6; s++) cand[s] = false;
        } else {
            uint32_t m = 0;
#pragma unroll
            for (int s = 0; s < 6; s++)
                if (cand[s]) {
                    __stcg(&lab[nb[s]], mylab);
                    m = max(m, l[s]);
                }
            if (rstar != NONE32 && m > cur) atomicMax(&S.mx, m);
        }
        headc += rstar == NONE32 ? k : min(k, rstar + 1u);
#pragma unroll
        for (int s = 0; s < 6; s++) nb[s] |= (uint32_t)(s + 1) << 29;
        block_append<6>(S, cand, l, nb, queue, lvl_qstart, ltail, lnz, lo, cur, tailc, mx);
        if (rstar != NONE32) {
            intr++;
            if (tid == 0) {
                lhead[cur - lo] = headc;
                ltail[cur - lo] = tailc;
                if (headc != tailc)
                    lnz[(cur - lo) >> 5] |= 1u << ((cur - lo) & 31);
                else
                    lnz[(cur - lo) >> 5] &= ~(1u << ((cur - lo) & 31));
            }
            cur = mx;
            __syncthreads();
            headc = lhead[cur - lo];
            tailc = ltail[cur - lo];
            qs = __ldg(&lvl_qstart[cur]);
        } else
            __syncthreads();
    }
    if (tid == 0 && stats) {
        atomicAdd(&stats[0], steps);
        atomicAdd(&stats[1], intr);
        atomicMax(&stats[2], steps);
    }
}

// ------------------------------------------------------------------ flood v2 (2-D tiles up to 2^17 pixels)
// Same step semantics as k_flood, but the per-pixel state the hot loop touches lives on chip: the
// "in mask, not labelled yet" bitmap of the tile and a small claim table sit in shared memory (no global
// atomics, nothing to undo when a step is truncated), queue entries carry (label << 17 | pixel), levels
// are u16.  Global traffic per flooded pixel: one queue entry written + read, one u16 level read.
static constexpr int F2_WARPS = 1;
static constexpr int F2_HASH = 256;
static constexpr uint32_t F2_PIXMASK = 0x1FFFFu;
static constexpr int F2_MAXPIX = 1 << 17;

// GAVAIL: the "floodable" bitmap stays in global memory (L2) instead of shared memory: 1 KB of shared memory per tile, so
// every tile of a batch is resident at once (no second wave) and more warps hide the load latencies.
// SLEV (with GAVAIL): the tails of the tile's level FIFOs and a bitmap of the levels that hold entries live in shared
// memory (levcap levels per tile), so an append costs no L2 round trip and the next level is found by bit scans.
// FR: inputs of the fused front end (front2d.cuh): the tile's seeds are ready-made queue entries seedent[t.base + k]
// (k < t_nseeds[tile]), its levels occupy the fixed slots [tile * levtab, tile * levtab + t_nlev[tile]).
template <bool GAVAIL, bool SLEV, bool FR = false>
__global__ void __launch_bounds__(32 * F2_WARPS) k_flood2(const Tile *__restrict__ tiles, int ntiles,
                                                          const uint32_t *__restrict__ lab_all, const uint16_t *__restrict__ lv16_all,
                                                          uint32_t *__restrict__ availw_all, uint32_t *__restrict__ queue,
                                                          const uint32_t *__restrict__ lvl_qstart, uint32_t *__restrict__ lvl_head,
                                                          uint32_t *__restrict__ lvl_tail, const uint32_t *__restrict__ tile_lvl,
                                                          const uint32_t *__restrict__ seedlist, const uint32_t *__restrict__ tile_seed,
                                                          int nwords_max, int levcap, uint32_t *__restrict__ stats,
                                                          const uint32_t *__restrict__ seedent = nullptr,
                                                          const uint32_t *__restrict__ t_nlev = nullptr,
                                                          const uint32_t *__restrict__ t_nseeds = nullptr, int levtab = 0) {
    extern __shared__ uint32_t f2_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wid = blockIdx.x * F2_WARPS + warp;
    if (wid >= ntiles) return;
    const Tile t = tiles[wid];
    uint32_t *avail = GAVAIL ? availw_all + (t.base >> 5) : f2_smem + (size_t)warp * (nwords_max + F2_HASH);
    const int lvwords = (levcap + 31) >> 5;
    uint32_t *claim = GAVAIL ? f2_smem + (size_t)warp * (F2_HASH + (SLEV ? levcap + lvwords : 0)) : avail + nwords_max;
    uint32_t *stail = claim + F2_HASH, *snz = stail + levcap;   // SLEV only
    const uint16_t *lv16 = lv16_all + t.base;
    const int W = t.W, H = t.H;
    const int npix = H * W, nwords = (npix + 31) >> 5;
    const uint32_t lo = FR ? (uint32_t)wid * (uint32_t)levtab : tile_lvl[wid], hi = FR ? lo + t_nlev[wid] : tile_lvl[wid + 1];
    if (lo == hi) return;
    const uint32_t sb = FR ? 0u : tile_seed[wid], se = FR ? t_nseeds[wid] : tile_seed[wid + 1];
    if (sb == se) return;
    if (!GAVAIL) {
        for (int w = lane; w < nwords; w += 32) avail[w] = availw_all[(t.base >> 5) + w];
        __syncwarp();
    }
    if (SLEV) {
        for (int j = lane; j < levcap + lvwords; j += 32) stail[j] = 0;
        __syncwarp();
    }
    uint32_t steps = 0, intr = 0;
    uint32_t cur = lo, tailc = 0, headc = 0;
    {
        uint32_t maxr = lo, dummy_tail = 0;
        for (uint32_t s0 = sb; s0 < se; s0 += 32) {
            bool v[1];
            uint32_t l[1], ent[1];
            v[0] = s0 + lane < se;
            const uint32_t e0 = (FR && v[0]) ? seedent[t.base + s0 + lane] : 0u;
            uint32_t px = FR ? (e0 & F2_PIXMASK) : (v[0] ? seedlist[s0 + lane] : 0);
            l[0] = v[0] ? lo + lv16[px] : 0;
            ent[0] = FR ? e0 : (v[0] ? ((lab_all[t.base + px] << 17) | px) : 0);
            if (v[0]) maxr = max(maxr, l[0]);
            warp_append<1, SLEV>(v, l, ent, queue, lvl_qstart, lvl_tail, NONE32, dummy_tail, lane, stail, snz, lo);
            __syncwarp();
        }
        cur = __reduce_max_sync(FULL, maxr);
        headc = 0;
        tailc = SLEV ? stail[cur - lo] : __ldcg(&lvl_tail[cur]);
    }
    uint32_t qs = lvl_qstart[cur];
    uint32_t pre_level = NONE32, pre_head = 0, pre_tail = 0, pre_entry = 0;
    for (;;) {
        if (headc == tailc) {
            bool found = false;
            if (SLEV) {
                // level exhausted: clear its bit, the next level is the highest set bit below it
                if (lane == 0) {
                    __stcg(&lvl_head[cur], headc);
                    stail[cur - lo] = tailc;
                    snz[(cur - lo) >> 5] &= ~(1u << ((cur - lo) & 31));
                }
                __syncwarp();
                const int r = (int)(cur - lo);
                for (int wtop = (r - 1) >> 5; r > 0 && wtop >= 0; wtop -= 32) {
                    const int wi = wtop - lane;
                    uint32_t wv = wi >= 0 ? snz[wi] : 0u;
                    if (wi == ((r - 1) >> 5) && (r & 31)) wv &= (1u << (r & 31)) - 1u;
                    const unsigned bb = __ballot_sync(FULL, wv != 0);
                    if (bb) {
                        const int src = __ffs(bb) - 1;
                        const uint32_t wsel = __shfl_sync(FULL, wv, src);
                        cur = lo + (uint32_t)((wtop - src) * 32 + 31 - __clz(wsel));
                        found = true;
                        break;
                    }
                }
                if (!found) break;
                headc = __ldcg(&lvl_head[cur]);
                tailc = stail[cur - lo];
            } else {
                if (lane == 0) {
                    __stcg(&lvl_head[cur], headc);
                    __stcg(&lvl_tail[cur], tailc);
                }
                __syncwarp();
                long long r0 = (long long)cur - 1;
                while (r0 >= (long long)lo) {
                    long long r = r0 - lane;
                    bool ne = false;
                    if (r >= (long long)lo) ne = __ldcg(&lvl_tail[r]) != __ldcg(&lvl_head[r]);
                    unsigned b = __ballot_sync(FULL, ne);
                    if (b) {
                        cur = (uint32_t)(r0 - (__ffs(b) - 1));
                        found = true;
                        break;
                    }
                    r0 -= 32;
                }
                if (!found) break;
                headc = __ldcg(&lvl_head[cur]);
                tailc = __ldcg(&lvl_tail[cur]);
            }
            qs = __ldg(&lvl_qstart[cur]);
            continue;
        }
        steps++;
        const uint32_t k = min(32u, tailc - headc);
        const bool act = (uint32_t)lane < k;
        // entries that were already queued when the previous step ended were fetched during that step
        uint32_t entry = 0;
        if (act) entry = (pre_level == cur && pre_head == headc && headc + lane < pre_tail) ? pre_entry : __ldcg(&queue[qs + headc + lane]);
        const uint32_t p = entry & F2_PIXMASK, mylab = entry >> 17;
        const int y = (int)fdiv(p, t.fW), x = (int)p - y * W;
        // neighbour order of skimage (connectivity 1, 2-D): -y, -x, +x, +y
        uint32_t nb[4];
        bool cand[4], pend[4];
        uint32_t hs[4];
        uint16_t lraw[4];
        nb[0] = (act && y > 0) ? p - W : NONE32;
        nb[1] = (act && x > 0) ? p - 1 : NONE32;
        nb[2] = (act && x + 1 < W) ? p + 1 : NONE32;
        nb[3] = (act && y + 1 < H) ? p + W : NONE32;
#pragma unroll
        for (int j = 0; j < F2_HASH / 32; j++) claim[lane + 32 * j] = NONE32;
#pragma unroll
        for (int s = 0; s < 4; s++) {
            cand[s] = nb[s] != NONE32 && (((GAVAIL ? __ldcg(&avail[nb[s] >> 5]) : avail[nb[s] >> 5]) >> (nb[s] & 31)) & 1u);
            pend[s] = cand[s];
            hs[s] = (nb[s] * 2654435761u) >> 24;
            lraw[s] = cand[s] ? lv16[nb[s]] : (uint16_t)0;   // issued early: the claim resolution hides the latency
        }
        __syncwarp();
        // the lowest (lane, slot) key wins every contested pixel
        for (;;) {
#pragma unroll
            for (int s = 0; s < 4; s++)
                if (pend[s]) atomicMin(&claim[hs[s]], ((uint32_t)(lane * 4 + s) << 17) | nb[s]);
            __syncwarp();
            bool again = false;
#pragma unroll
            for (int s = 0; s < 4; s++)
                if (pend[s]) {
                    uint32_t v = claim[hs[s]];
                    if ((v & F2_PIXMASK) == nb[s]) {
                        cand[s] = (v >> 17) == (uint32_t)(lane * 4 + s);
                        pend[s] = false;
                    } else {
                        hs[s] = (hs[s] + 1) & (F2_HASH - 1);
                        again = true;
                    }
                }
            if (!__any_sync(FULL, again)) break;
        }
        uint32_t l[4], ent[4];
        bool up = false;
#pragma unroll
        for (int s = 0; s < 4; s++) {
            l[s] = 0;
            ent[s] = 0;
            if (cand[s]) {
                l[s] = lo + lraw[s];
                ent[s] = (mylab << 17) | nb[s];
                up |= l[s] > cur;
            }
        }
        const unsigned ball = __ballot_sync(FULL, up);
        const int rstar = ball ? __ffs(ball) - 1 : 31;
#pragma unroll
        for (int s = 0; s < 4; s++) {
            if (lane > rstar) cand[s] = false;
            if (cand[s]) atomicAnd(&avail[nb[s] >> 5], ~(1u << (nb[s] & 31)));
        }
        headc += min(k, (uint32_t)rstar + 1u);
        // prefetch the next batch of this level (what is queued so far) behind the append
        pre_level = cur, pre_head = headc, pre_tail = tailc;
        pre_entry = headc + lane < tailc ? __ldcg(&queue[qs + headc + lane]) : 0;
        warp_append<4, SLEV>(cand, l, ent, queue, lvl_qstart, lvl_tail, cur, tailc, lane, stail, snz, lo);
        __syncwarp();
        if (ball) {
            intr++;
            uint32_t mx = 0;
#pragma unroll
            for (int s = 0; s < 4; s++)
                if (cand[s]) mx = max(mx, l[s]);
            mx = __reduce_max_sync(FULL, mx);
            if (lane == 0) {
                __stcg(&lvl_head[cur], headc);
                if (SLEV) {
                    stail[cur - lo] = tailc;
                    if (headc != tailc)
                        snz[(cur - lo) >> 5] |= 1u << ((cur - lo) & 31);
                    else
                        snz[(cur - lo) >> 5] &= ~(1u << ((cur - lo) & 31));
                } else
                    __stcg(&lvl_tail[cur], tailc);
            }
            __syncwarp();
            cur = mx;
            headc = __ldcg(&lvl_head[cur]);
            tailc = SLEV ? stail[cur - lo] : __ldcg(&lvl_tail[cur]);
            qs = __ldg(&lvl_qstart[cur]);
        }
    }
    if (lane == 0 && stats) {
        atomicAdd(&stats[0], steps);
        atomicAdd(&stats[1], intr);
        atomicMax(&stats[2], steps);
    }
}

// labels of flood v2: every queue entry (seed or flooded pixel) carries its label
__global__ void __launch_bounds__(256) k_scatter_labels(const Tile *__restrict__ tiles, const uint32_t *__restrict__ tile_q,
                                                        const uint32_t *__restrict__ queue, uint32_t *__restrict__ lab) {
    const Tile t = tiles[blockIdx.y];
    const uint32_t qb = tile_q[blockIdx.y], qe = tile_q[blockIdx.y + 1];
    for (uint32_t q = qb + blockIdx.x * blockDim.x + threadIdx.x; q < qe; q += gridDim.x * blockDim.x) {
        uint32_t e = queue[q];
        if (e != NONE32) lab[t.base + (e & F2_PIXMASK)] = e >> 17;
    }
}

// the same through shared memory when a tile's 16-bit labels fit there (one CTA per tile): the scattered writes stay on
// chip and the label plane leaves with coalesced stores; mask pixels the flood did not reach get 0 instead of UNLAB (every
// consumer treats both as "no fragment")
static constexpr int SCAT_NT = 1024;
__global__ void __launch_bounds__(SCAT_NT) k_scatter_labels_tile(const Tile *__restrict__ tiles, const uint32_t *__restrict__ tile_q,
                                                                 const uint32_t *__restrict__ queue, uint32_t *__restrict__ lab) {
    extern __shared__ uint16_t sc_lab[];
    const Tile t = tiles[blockIdx.x];
    const int npix = t.H * t.W;
    for (int i = threadIdx.x; i < (npix + 1) / 2; i += SCAT_NT) ((uint32_t *)sc_lab)[i] = 0u;
    __syncthreads();
    const uint32_t qb = tile_q[blockIdx.x], qe = tile_q[blockIdx.x + 1];
    for (uint32_t q = qb + threadIdx.x; q < qe; q += SCAT_NT) {
        const uint32_t e = __ldcs(&queue[q]);
        if (e != NONE32) sc_lab[e & F2_PIXMASK] = (uint16_t)(e >> 17);
    }
    __syncthreads();
    uint32_t *out = lab + t.base;
    for (int i = threadIdx.x; i < npix; i += SCAT_NT) out[i] = sc_lab[i];
}

// ------------------------------------------------------------------ fragment statistics
// One table entry per watershed fragment (index = fbase[tile] + label - 1): affinity sum + voxel count over the
// whole read-ROI tile (filter_avg_fragments / remove_small_objects see the uncropped array,
// watershed_frags.py:148-156,188-192), the smallest write-order index of its voxels inside the write ROI, and
// whether it also has voxels outside the write ROI ("crossing": only such a fragment can fall apart when cropped).
static constexpr uint8_t FF_CROSS = 1, FF_KEEP = 2;

__device__ __forceinline__ long long tile_widx(const Tile &t, int z, int y, int x) {
    return t.wbase + ((long long)(z - t.wz) * t.wH + (y - t.wy)) * t.wW + (x - t.wx);
}

// label planes: u32 (unfused chain; may hold claim keys / UNLAB = no fragment) or u16 (fused front end; 0 = no fragment)
__device__ __forceinline__ uint32_t lab_get(const uint32_t *lab, long long i) {
    const uint32_t l = lab[i];
    return l >= CLAIM ? 0u : l;
}
__device__ __forceinline__ uint32_t lab_get(const uint16_t *lab, long long i) { return lab[i]; }

template <typename T, typename LT>
__global__ void __launch_bounds__(256) k_fragstats(const Tile *__restrict__ tiles, AffView A, const LT *__restrict__ lab,
                                                   const uint32_t *__restrict__ fbase, int need_stats, int all_cross,
                                                   typename AffOps<T>::acc_t *__restrict__ fsum, uint32_t *__restrict__ fcnt,
                                                   uint32_t *__restrict__ fmin, uint8_t *__restrict__ fflag) {
    const Tile t = tiles[blockIdx.y];
    const int W = t.W, H = t.H;
    const long long HW = (long long)H * W, npix = (long long)t.D * HW;
    const size_t nvol = (size_t)A.Zw * A.Y * A.X;
    const T *a = (const T *)A.p;
    const uint32_t fb = fbase[blockIdx.y];
    for (long long i0 = (long long)blockIdx.x * blockDim.x; i0 < npix; i0 += (long long)gridDim.x * blockDim.x) {
        long long i = i0 + threadIdx.x;
        uint32_t l = 0, w = NONE32;
        unsigned outside = 0;
        typename AffOps<T>::acc_t val = 0;
        if (i < npix) {
            l = lab_get(lab, t.base + i);
            if (l) {
                // a labelled pixel is inside the mask, hence inside the volume and not masked out
                int x, y, z;
                unravel3f((uint32_t)i, W, H, t.fW, t.fH, x, y, z);
                if (z >= t.wz && z < t.wz + t.wD && y >= t.wy && y < t.wy + t.wH && x >= t.wx && x < t.wx + t.wW)
                    w = (uint32_t)tile_widx(t, z, y, x);
                else
                    outside = 1;
                if (need_stats) {
                    // a watershed fragment lies inside the mask, hence inside the volume; given labels (bs_stage1_from_labels:
                    // mutex-watershed fragments) also cover the zero-filled margin and masked-out voxels, whose affinity is 0
                    const int gz = t.gz + z, gy = t.gy + y, gx = t.gx + x;
                    if (gz >= A.z0 && gz < A.z0 + A.Zw && gy >= 0 && gy < A.Y && gx >= 0 && gx < A.X) {
                        size_t gi = ((size_t)(gz - A.z0) * A.Y + gy) * A.X + gx;
                        if (!A.mask || A.mask[gi] > 0) val = AffOps<T>::value(a, nvol, gi);
                    }
                }
            }
        }
        unsigned act = __ballot_sync(FULL, l != 0);
        if (l) {
            unsigned peers = __match_any_sync(act, l);
            const bool leader = (int)(threadIdx.x & 31) == __ffs(peers) - 1;
            const size_t fi = (size_t)fb + l - 1;
            uint32_t wmin = __reduce_min_sync(peers, w);
            unsigned anyout = __reduce_or_sync(peers, outside);
            if constexpr (sizeof(T) == 1) {
                unsigned sm = __reduce_add_sync(peers, (unsigned)val);
                if (leader && need_stats) atomicAdd(&fsum[fi], (unsigned long long)sm);
            } else {
                if (need_stats) atomicAdd(&fsum[fi], val);
            }
            if (leader) {
                atomicAdd(&fcnt[fi], (uint32_t)__popc(peers));
                if (wmin != NONE32) atomicMin(&fmin[fi], wmin);
                // given fragments (mutex watershed with long-range edges) need not be connected: all go through the relabel
                if (anyout || all_cross) fflag[fi] = FF_CROSS;   // every writer stores the same value
            }
        }
    }
}

// ------------------------------------------------------------------ crop + full-connectivity relabel
template <typename ACC>
__device__ __forceinline__ bool frag_keep(ACC sum, uint32_t cnt, double ff, int rd, bool is_u8) {
    if (ff > 0.0) {
        double mean = is_u8 ? ((double)sum / 765.0) / (double)cnt : (double)sum / (double)cnt;
        if (mean < ff) return false;            // watershed_frags.py:153
    }
    if (rd > 0 && cnt < (uint32_t)rd) return false;  // remove_small_objects(min_size=rd)
    return true;
}

// one decision per fragment
template <typename ACC>
__global__ void __launch_bounds__(256) k_frag_decide(const ACC *__restrict__ fsum, const uint32_t *__restrict__ fcnt,
                                                     uint8_t *__restrict__ fflag, size_t n, double ff, int rd, int is_u8) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t c = fcnt[i];
    if (c && frag_keep<ACC>(fsum[i], c, ff, rd, is_u8 != 0)) fflag[i] |= FF_KEEP;
}

// Write-ROI voxels of kept crossing fragments take part in a pixel-level union-find (cpar, tile-local indices):
// parent = start of the voxel's row run of equal labels inside its warp chunk (pre-linked runs keep chains
// short), NONE32 for every other write-ROI voxel.  Fragments entirely inside the write ROI stay connected.
// xbits: one bit per write-ROI voxel (batch write order): the voxel takes part in the union-find.  Only those voxels'
// cpar entries are ever written or read.
template <typename LT>
__global__ void __launch_bounds__(256) k_crop_init(const Tile *__restrict__ tiles, const LT *__restrict__ lab,
                                                   const uint32_t *__restrict__ fbase, const uint8_t *__restrict__ fflag,
                                                   uint32_t *__restrict__ cpar, uint32_t *__restrict__ xbits) {
    const Tile t = tiles[blockIdx.y];
    const int W = t.W, H = t.H;
    const long long nw = (long long)t.wD * t.wH * t.wW;
    const int lane = threadIdx.x & 31;
    const uint32_t fb = fbase[blockIdx.y];
    auto part_label = [&](long long i) -> uint32_t {
        uint32_t l = lab_get(lab, t.base + i);
        if (l && (fflag[(size_t)fb + l - 1] & (FF_CROSS | FF_KEEP)) == (FF_CROSS | FF_KEEP)) return l;
        return 0;
    };
    for (long long k0 = (long long)blockIdx.x * blockDim.x; k0 < nw; k0 += (long long)gridDim.x * blockDim.x) {
        long long kk = k0 + threadIdx.x;
        uint32_t l = 0;
        long long i = 0;
        int x = 0, y = 0, z = 0;
        if (kk < nw) {
            unravel3f((uint32_t)kk, t.wW, t.wH, t.fwW, t.fwH, x, y, z);
            i = ((long long)(z + t.wz) * H + (y + t.wy)) * W + (x + t.wx);
            l = part_label(i);
        }
        uint32_t ll = __shfl_up_sync(FULL, l, 1);
        if (lane == 0) ll = (l && x > 0) ? part_label(i - 1) : 0;
        bool sl = l && x > 0 && ll == l;
        unsigned startbits = __ballot_sync(FULL, l && !sl);
        if (kk < nw && l) {
            unsigned m = startbits & (FULL >> (31 - lane));
            cpar[t.base + i] = m ? (uint32_t)(i - (lane - (31 - __clz(m)))) : (uint32_t)(i - lane);
        }
        const unsigned part = __ballot_sync(FULL, l != 0);
        if (lane == 0 && part) {
            const unsigned long long w0 = (unsigned long long)(t.wbase + k0 + (threadIdx.x & ~31));
            const unsigned sh = (unsigned)(w0 & 31);
            atomicOr(&xbits[w0 >> 5], part << sh);
            if (sh && (part >> (32 - sh))) atomicOr(&xbits[(w0 >> 5) + 1], part >> (32 - sh));
        }
    }
}

// union with the raster-preceding neighbours of the full (8 / 26) neighbourhood inside the write ROI that
// carry the same label; links implied by row adjacency of equal labels are skipped
template <typename LT>
__global__ void __launch_bounds__(256) k_crop_union(const Tile *__restrict__ tiles, const LT *__restrict__ lab,
                                                    uint32_t *__restrict__ cpar, const uint32_t *__restrict__ xbits) {
    const Tile t = tiles[blockIdx.y];
    const int W = t.W, H = t.H;
    const long long HW = (long long)H * W;
    const long long nw = (long long)t.wD * t.wH * t.wW;
    uint32_t *pp = cpar + t.base;
    const LT *ll = lab + t.base;
    for (long long kk = (long long)blockIdx.x * blockDim.x + threadIdx.x; kk < nw; kk += (long long)gridDim.x * blockDim.x) {
        const unsigned long long wq = (unsigned long long)(t.wbase + kk);
        if (!((xbits[wq >> 5] >> (wq & 31)) & 1u)) continue;
        int x, y, z;
        unravel3f((uint32_t)kk, t.wW, t.wH, t.fwW, t.fwH, x, y, z);
        const long long i = ((long long)(z + t.wz) * H + (y + t.wy)) * W + (x + t.wx);
        const uint32_t l = ll[i];
        // every neighbour tested below lies inside the write ROI, where "same label" means "same fragment" and hence
        // "takes part too" (crossing / kept are properties of the fragment)
        auto same = [&](long long j) -> bool { return (uint32_t)ll[j] == l; };
        const bool left = x > 0 && same(i - 1);
        // row link across a warp-chunk boundary (inside a chunk k_crop_init linked the run already)
        if (left && (kk & 31) == 0) uf_union(pp, (uint32_t)i, (uint32_t)(i - 1));
        if (y > 0) {
            const bool up = same(i - W);
            if (up) {
                if (!(left && same(i - W - 1))) uf_union(pp, (uint32_t)i, (uint32_t)(i - W));
            } else {
                if (x > 0 && !left && same(i - W - 1)) uf_union(pp, (uint32_t)i, (uint32_t)(i - W - 1));
                if (x + 1 < t.wW && same(i - W + 1) && !same(i + 1)) uf_union(pp, (uint32_t)i, (uint32_t)(i - W + 1));
            }
        }
        if (z > 0) {
            const long long b = i - HW;
            if (same(b)) {
                if (!(left && same(b - 1))) uf_union(pp, (uint32_t)i, (uint32_t)b);
            } else {
                for (int dy = -1; dy <= 1; dy++)
                    for (int dx = -1; dx <= 1; dx++) {
                        if (dy == 0 && dx == 0) continue;
                        int yy = y + dy, xx = x + dx;
                        if (yy < 0 || yy >= t.wH || xx < 0 || xx >= t.wW) continue;
                        long long j = b + (long long)dy * W + dx;
                        if (same(j)) uf_union(pp, (uint32_t)i, (uint32_t)j);
                    }
            }
        }
    }
}

// one bit per write-ROI voxel (batch write order): set for the first voxel of every output fragment
__global__ void __launch_bounds__(256) k_root_bits_pix(const Tile *__restrict__ tiles, const uint32_t *__restrict__ cpar,
                                                       const uint32_t *__restrict__ xbits, uint32_t *__restrict__ bits) {
    const Tile t = tiles[blockIdx.y];
    const int W = t.W, H = t.H;
    const long long nw = (long long)t.wD * t.wH * t.wW;
    for (long long kk = (long long)blockIdx.x * blockDim.x + threadIdx.x; kk < nw; kk += (long long)gridDim.x * blockDim.x) {
        const uint32_t w = (uint32_t)(t.wbase + kk);
        if (!((xbits[w >> 5] >> (w & 31)) & 1u)) continue;
        int x, y, z;
        unravel3f((uint32_t)kk, t.wW, t.wH, t.fwW, t.fwH, x, y, z);
        const long long i = ((long long)(z + t.wz) * H + (y + t.wy)) * W + (x + t.wx);
        if (cpar[t.base + i] == (uint32_t)i) atomicOr(&bits[w >> 5], 1u << (w & 31));
    }
}

__global__ void __launch_bounds__(256) k_root_bits_frag(const uint32_t *__restrict__ fcnt, const uint8_t *__restrict__ fflag,
                                                        const uint32_t *__restrict__ fmin, size_t n, uint32_t *__restrict__ bits) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (fcnt[i] && (fflag[i] & (FF_CROSS | FF_KEEP)) == FF_KEEP) {
        const uint32_t w = fmin[i];
        atomicOr(&bits[w >> 5], 1u << (w & 31));
    }
}

__global__ void k_popc_words(const uint32_t *__restrict__ bits, uint32_t *__restrict__ cnt, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cnt[i] = __popc(bits[i]);
}

struct BlkDev {
    long long block_id;
    long long wbase;       // batch write-order base of the block
    int wo[3], ws[3];
    int plan_index;
    int pad_;
};

__global__ void k_blk_first(const BlkDev *__restrict__ blks, int nblk, const uint32_t *__restrict__ bits,
                            const uint32_t *__restrict__ wscan, uint32_t *__restrict__ blk_first) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nblk) blk_first[i] = bit_rank(bits, wscan, (uint32_t)blks[i].wbase);
}

// ids (raster order of each fragment's first voxel, + block_id * prod(block_size)), uint64 output, node
// statistics.  One CTA column per tile (blockIdx.y).
template <typename LT>
__global__ void __launch_bounds__(256) k_finalize(const Tile *__restrict__ tiles, const BlkDev *__restrict__ blks,
                                                  const LT *__restrict__ lab, const uint32_t *__restrict__ cpar,
                                                  const uint32_t *__restrict__ fbase, const uint8_t *__restrict__ fflag,
                                                  const uint32_t *__restrict__ fmin, const uint32_t *__restrict__ bits,
                                                  const uint32_t *__restrict__ wscan, const uint32_t *__restrict__ blk_first,
                                                  long long nvox_block, int roi_oz, int roi_oy, int roi_ox, int roi_Y, int roi_X,
                                                  uint64_t *__restrict__ frags, uint32_t *__restrict__ ncnt,
                                                  unsigned long long *__restrict__ nsum) {
    const Tile t = tiles[blockIdx.y];
    const BlkDev b = blks[t.block];
    const int W = t.W, H = t.H;
    const long long nw = (long long)t.wD * t.wH * t.wW;
    // a tile's write region is the block's write ROI (3-D mode) or one z plane of it (xy mode): block coordinates are the
    // tile's write coordinates plus the plane number
    const int zoff = (int)((t.wbase - b.wbase) / ((long long)t.wH * t.wW));
    const uint32_t first = blk_first[t.block];
    const uint32_t fb = fbase[blockIdx.y];
    const uint32_t *pp = cpar + t.base;
    for (long long k0 = (long long)blockIdx.x * blockDim.x; k0 < nw; k0 += (long long)gridDim.x * blockDim.x) {
        long long kk = k0 + threadIdx.x;
        uint32_t node = NONE32;
        int x = 0, y = 0, z = 0;
        if (kk < nw) {
            int tx, ty, tz;
            unravel3f((uint32_t)kk, t.wW, t.wH, t.fwW, t.fwH, tx, ty, tz);
            const long long i = ((long long)(tz + t.wz) * H + (ty + t.wy)) * W + (tx + t.wx);
            x = tx, y = ty, z = tz + zoff;
            const uint32_t l = lab_get(lab, t.base + i);
            uint64_t id = 0;
            if (l) {
                const size_t fi = (size_t)fb + l - 1;
                const uint8_t fl = fflag[fi];
                if (fl & FF_KEEP) {
                    uint32_t rw;
                    if (fl & FF_CROSS) {
                        uint32_t root = uf_find(pp, (uint32_t)i);
                        int rx, ry, rz;
                        unravel3f(root, W, H, t.fW, t.fH, rx, ry, rz);
                        rw = (uint32_t)tile_widx(t, rz, ry, rx);
                    } else {
                        rw = fmin[fi];
                    }
                    node = bit_rank(bits, wscan, rw);
                    id = (uint64_t)(node - first + 1) + (uint64_t)b.block_id * (uint64_t)nvox_block;
                }
            }
            size_t o = ((size_t)(b.wo[0] + z - roi_oz) * roi_Y + (b.wo[1] + y - roi_oy)) * roi_X + (b.wo[2] + x - roi_ox);
            frags[o] = id;
        }
        unsigned act = __ballot_sync(FULL, node != NONE32);
        if (node != NONE32) {
            unsigned peers = __match_any_sync(act, node);
            int leader = __ffs(peers) - 1;
            unsigned sz = __reduce_add_sync(peers, (unsigned)z);
            unsigned sy = __reduce_add_sync(peers, (unsigned)y);
            unsigned sx = __reduce_add_sync(peers, (unsigned)x);
            if ((threadIdx.x & 31) == leader) {
                atomicAdd(&ncnt[node], (uint32_t)__popc(peers));
                atomicAdd(&nsum[3 * (size_t)node + 0], (unsigned long long)sz);
                atomicAdd(&nsum[3 * (size_t)node + 1], (unsigned long long)sy);
                atomicAdd(&nsum[3 * (size_t)node + 2], (unsigned long long)sx);
            }
        }
    }
}

// node table: id, position = write offset + trunc(centre of mass), size   (watershed_frags.py:230-246)
__global__ void k_nodes(const BlkDev *__restrict__ blks, int nblk, const uint32_t *__restrict__ blk_first, uint32_t n_nodes,
                        const uint32_t *__restrict__ ncnt, const unsigned long long *__restrict__ nsum, long long nvox_block,
                        uint64_t *__restrict__ ids, int32_t *__restrict__ pos, uint32_t *__restrict__ sizes) {
    uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_nodes) return;
    // block of node n: last block with blk_first <= n (blocks are in ascending order)
    int lo = 0, hi = nblk - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (blk_first[mid] <= n)
            lo = mid;
        else
            hi = mid - 1;
    }
    const BlkDev b = blks[lo];
    uint32_t c = ncnt[n];
    ids[n] = (uint64_t)(n - blk_first[lo] + 1) + (uint64_t)b.block_id * (uint64_t)nvox_block;
    sizes[n] = c;
    for (int d = 0; d < 3; d++) pos[3 * (size_t)n + d] = b.wo[d] + (int32_t)(nsum[3 * (size_t)n + d] / c);
}

// ------------------------------------------------------------------ host driver
struct TileDims {
    int ntiles, maxD, maxH, maxW;
    long long maxpix;
};

static dim3 pixel_grid(const TileDims &td) {
    return dim3((unsigned)std::min<long long>(std::max<long long>((td.maxpix + PIX_PER_CTA - 1) / PIX_PER_CTA, 1), 2048), td.ntiles);
}

// exact squared EDT of the mask whose row distances are in g -> out (tmp: in-plane result of 3-D tiles)
static int launch_edt(const Tile *dt, const TileDims &td, bool three_d, const uint16_t *g, uint32_t *tmp, uint32_t *out,
                      uint32_t *tilemax, cudaStream_t s) {
    const dim3 grid = pixel_grid(td);
    const size_t strip_smem = (size_t)td.maxH * CS_W * 4;
    const bool use_strip = strip_smem <= 96 * 1024;
    const dim3 grid_strip((unsigned)std::min<long long>((long long)td.maxD * ((td.maxW + CS_W - 1) / CS_W), 8192), td.ntiles);
    if (use_strip) {
        static std::atomic<unsigned long long> attr_strip(0);   // one bit per device: function attributes are per device
        if (!dev_once(attr_strip)) {
            BS_CUDA(cudaFuncSetAttribute(k_coldist_strip, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        }
    }
    uint32_t *plane_out = three_d ? tmp : out;
    if (use_strip)
        BS_LAUNCH(k_coldist_strip, grid_strip, 256, strip_smem, s, dt, g, plane_out, tilemax);
    else
        BS_LAUNCH(k_coldist, grid, 256, 0, s, dt, g, plane_out, tilemax);
    if (three_d) BS_LAUNCH(k_zdist, grid, 256, 0, s, dt, tmp, out, tilemax);
    return BS_OK;
}

// seeds = (maximum_filter(d2, msd) == d2) & msk: parent array (own index / NONE32) + seed bitmap
static int launch_seeds(const Tile *dt, const TileDims &td, bool three_d, int msd, const uint32_t *d2, const uint8_t *msk,
                        uint32_t *tmpA, uint32_t *tmpB, uint32_t *par, uint32_t *sbits, cudaStream_t s) {
    const dim3 grid = pixel_grid(td);
    const size_t mf_smem = ((size_t)(MF_TH + msd - 1) * (MF_TW + msd - 1) + (size_t)(MF_TH + msd - 1) * MF_TW) * 4;
    const bool use_mf = mf_smem <= 96 * 1024;
    const dim3 grid_mf((unsigned)std::min<long long>(
                           (long long)td.maxD * ((td.maxW + MF_TW - 1) / MF_TW) * ((td.maxH + MF_TH - 1) / MF_TH), 8192),
                       td.ntiles);
    if (use_mf) {
        static std::atomic<unsigned long long> attr_mf(0);   // one bit per device: function attributes are per device
        if (!dev_once(attr_mf)) {
            BS_CUDA(cudaFuncSetAttribute(k_maxfilt_xy, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        }
    }
    if (!three_d) {
        if (msd == 10) {
            BS_LAUNCH(k_maxfilt_xy_t<10>, grid_mf, 256, 0, s, dt, d2, 1, msk, par, sbits);
        } else if (use_mf) {
            BS_LAUNCH(k_maxfilt_xy, grid_mf, 256, mf_smem, s, dt, d2, msd, 1, msk, par, sbits);
        } else {
            BS_LAUNCH(k_maxfilt, grid, 256, 0, s, dt, d2, tmpA, 2, msd, 0, nullptr, nullptr, nullptr, nullptr);
            BS_LAUNCH(k_maxfilt, grid, 256, 0, s, dt, tmpA, nullptr, 1, msd, 1, d2, msk, par, sbits);
        }
    } else {
        if (msd == 10) {
            BS_LAUNCH(k_maxfilt_xy_t<10>, grid_mf, 256, 0, s, dt, d2, 0, nullptr, tmpB, nullptr);
        } else if (use_mf) {
            BS_LAUNCH(k_maxfilt_xy, grid_mf, 256, mf_smem, s, dt, d2, msd, 0, nullptr, tmpB, nullptr);
        } else {
            BS_LAUNCH(k_maxfilt, grid, 256, 0, s, dt, d2, tmpA, 2, msd, 0, nullptr, nullptr, nullptr, nullptr);
            BS_LAUNCH(k_maxfilt, grid, 256, 0, s, dt, tmpA, tmpB, 1, msd, 0, nullptr, nullptr, nullptr, nullptr);
        }
        BS_LAUNCH(k_maxfilt, grid, 256, 0, s, dt, tmpB, nullptr, 0, msd, 1, d2, msk, par, sbits);
    }
    return BS_OK;
}

static int rowdist_grid_rows(const std::vector<Tile> &tiles) {
    int rows = 0;
    for (auto &t : tiles) rows = std::max(rows, t.D * t.H);
    return std::min(std::max((rows + 7) / 8, 1), 4096);
}

// 3-D tiles = read ROIs of the blocks of a batch: what the seed_eps and sigma pre-passes work on
struct PreTiles {
    std::vector<Tile> pt;
    std::vector<PreRef> refs;   // per block of the batch
    long long Ppre = 0;
    TileDims td;
    DevBuf d_pt;
};

static int build_pretiles(Plan &P, const std::vector<int> &bidx, PreTiles &R, cudaStream_t s) {
    long long maxpix = 0;
    for (size_t bi = 0; bi < bidx.size(); bi++) {
        const Blk &b = P.blocks[bidx[bi]];
        Tile t;
        t.gz = b.ro[0], t.gy = b.ro[1], t.gx = b.ro[2];
        t.D = b.rs[0], t.H = b.rs[1], t.W = b.rs[2];
        t.wz = t.wy = t.wx = 0;
        t.wD = t.D, t.wH = t.H, t.wW = t.W;
        t.block = (int)bi;
        t.ndim = 3;
        t.base = R.Ppre;
        t.wbase = 0;
        PreRef r;
        r.base = R.Ppre, r.oz = t.gz, r.oy = t.gy, r.ox = t.gx, r.H = t.H, r.W = t.W, r.D = t.D;
        r.block_id = b.block_id;
        R.refs.push_back(r);
        long long np = (long long)t.D * t.H * t.W;
        R.Ppre += np;
        maxpix = std::max(maxpix, np);
        t.set_divs();
        R.pt.push_back(t);
    }
    BS_ARG(R.Ppre < (1LL << 31), "stage1: seed_eps / sigma pre-pass too large for 32-bit tile indices (lower max_batch_voxels)");
    for (auto &t : R.pt) BS_ARG(t.W <= MAXW, "stage1: tile wider than 4096 voxels is not supported");
    R.td.ntiles = (int)R.pt.size(), R.td.maxpix = maxpix, R.td.maxD = R.td.maxH = R.td.maxW = 0;
    for (auto &t : R.pt)
        R.td.maxH = std::max(R.td.maxH, t.H), R.td.maxW = std::max(R.td.maxW, t.W), R.td.maxD = std::max(R.td.maxD, t.D);
    BS_TRY(R.d_pt.alloc(sizeof(Tile) * R.pt.size(), s));
    BS_CUDA(cudaMemcpyAsync(R.d_pt.p, R.pt.data(), sizeof(Tile) * R.pt.size(), cudaMemcpyHostToDevice, s));
    BS_CUDA(cudaStreamSynchronize(s));   // pt is a host-staged copy
    return BS_OK;
}

// seed_eps pre-pass (watershed_frags.py:131-139) on the 3-D read ROI of every block of the batch:
//   boundary_mask = mean(affs) > 0.5; seeds = (maximum_filter(EDT(mask), msd) == EDT(mask)) & mask;
//   D = EDT(seeds == 0)            -> D2 (exact squared distances; the sqrt is taken where the shift is applied)
template <typename T>
static int seed_distance_prepass(Plan &P, const PreTiles &R, AffView A, DevBuf &D2, cudaStream_t s) {
    const bs_ws_config &cfg = P.cfg;
    const long long Ppre = R.Ppre;
    const int np_tiles = (int)R.pt.size();
    const TileDims &td = R.td;
    const Tile *dt = R.d_pt.as<Tile>();
    DevBuf msk, g, d2, tmpA, tmpB, par, sb, flags, tmax;
    BS_TRY(msk.alloc(Ppre, s));
    BS_TRY(g.alloc(Ppre * 2, s));
    BS_TRY(d2.alloc(Ppre * 4, s));
    BS_TRY(tmpA.alloc(Ppre * 4, s));
    BS_TRY(tmpB.alloc(Ppre * 4, s));
    BS_TRY(par.alloc(Ppre * 4, s));
    BS_TRY(sb.alloc_zero(4 * ((size_t)Ppre / 32 + 2), s));
    BS_TRY(flags.alloc_zero(4 * (np_tiles + 1), s));
    BS_TRY(tmax.alloc_zero(4 * (np_tiles + 1), s));
    BS_TRY(D2.alloc(Ppre * 4, s));
    ShiftView none;
    memset(&none, 0, sizeof(none));
    const dim3 gr((unsigned)rowdist_grid_rows(R.pt), np_tiles);
    BS_LAUNCH((k_mask_rowdist<T, 0>), gr, 256, 0, s, dt, A, none, nullptr, msk.as<uint8_t>(), g.as<uint16_t>(), flags.as<uint32_t>());
    BS_TRY(launch_edt(dt, td, true, g.as<uint16_t>(), tmpA.as<uint32_t>(), d2.as<uint32_t>(), tmax.as<uint32_t>(), s));
    BS_TRY(launch_seeds(dt, td, true, cfg.min_seed_distance, d2.as<uint32_t>(), msk.as<uint8_t>(), tmpA.as<uint32_t>(),
                        tmpB.as<uint32_t>(), par.as<uint32_t>(), sb.as<uint32_t>(), s));
    BS_LAUNCH((k_mask_rowdist<T, 2>), gr, 256, 0, s, dt, A, none, par.as<uint32_t>(), nullptr, g.as<uint16_t>(), nullptr);
    BS_TRY(launch_edt(dt, td, true, g.as<uint16_t>(), tmpA.as<uint32_t>(), D2.as<uint32_t>(), tmax.as<uint32_t>(), s));
    return BS_OK;
}

// sigma pre-pass (watershed_frags.py:121-122): gaussian_filter of the block's (normalised, masked, zero-filled) read-ROI
// affinities in the dtype numpy holds them in (float64 for uint8 input, float32 for float32 input) -- gauss.cu
template <typename T, typename Out>
int gauss_tiles(const Tile *d_tiles, int ntiles, long long maxpix, const T *affs, const uint8_t *mask, int volZw, int volZ, int volY,
                int volX, int z0, const GaussWeights &gw, size_t cstride, Out *bufA, Out *bufB, Out **result, cudaStream_t s);

template <typename T>
static int gauss_prepass(Plan &P, const PreTiles &R, AffView A, DevBuf &bufA, DevBuf &bufB, DevBuf &wdev, const void **G, cudaStream_t s) {
    typedef typename std::conditional<sizeof(T) == 1, double, float>::type Out;
    const bs_ws_config &cfg = P.cfg;
    BS_TRY(bufA.alloc(sizeof(Out) * 3 * (size_t)R.Ppre, s));
    BS_TRY(bufB.alloc(sizeof(Out) * 3 * (size_t)R.Ppre, s));
    BS_TRY(wdev.alloc(sizeof(double) * 3 * BS_SIGMA_MAXW, s));
    BS_CUDA(cudaMemcpyAsync(wdev.p, cfg.sigma_w, sizeof(double) * 3 * BS_SIGMA_MAXW, cudaMemcpyHostToDevice, s));
    BS_CUDA(cudaStreamSynchronize(s));
    GaussWeights gw;
    for (int d = 0; d < 3; d++) {
        gw.radius[d] = cfg.sigma_radius[d];
        gw.w[d] = wdev.as<double>() + (size_t)d * BS_SIGMA_MAXW;
    }
    Out *res = nullptr;
    BS_TRY((gauss_tiles<T, Out>(R.d_pt.as<Tile>(), (int)R.pt.size(), R.td.maxpix, (const T *)A.p, A.mask, A.Zw, A.Z, A.Y, A.X, A.z0, gw,
                                (size_t)R.Ppre, bufA.as<Out>(), bufB.as<Out>(), &res, s)));
    *G = res;
    return BS_OK;
}

static int keep_debug(Plan &P, const char *name, DevBuf &buf, int elem, long long count) {
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

extern int g_debug;
extern int g_flood_version;

#include "front2d.cuh"

// what the front end (mask ... flood) leaves for the back end (fragment statistics ... ids)
struct S1Front {
    DevBuf lab, lv, tile_seed, totals, fstats;   // label plane, a scratch plane (cpar), fragment-table base per tile
    size_t nseeds = 0;                           // fragment table entries (labels are ranks of seed pixels when v2)
    bool v2 = false;
    bool lab16 = false;                          // the label plane holds 16-bit labels (fused front end)
};

// the unfused chain: one kernel per step, every intermediate plane in HBM.  Serves 3-D tiles, shifted affinities, tiles the
// fused front end cannot hold on chip, and debug runs (the intermediates are kept for bs_debug_fetch).
template <typename T>
static int stage1_front_unfused(Plan &P, const std::vector<int> &bidx, AffView A, const std::vector<Tile> &tiles, const Tile *dt,
                                long long P_pix, long long maxpix, dim3 grid, S1Front &F, cudaStream_t s) {
    const bs_ws_config &cfg = P.cfg;
    const bool xy = cfg.fragments_in_xy != 0;
    const int ntiles = (int)tiles.size();
    DevBuf msk, g, d2, tmpA, tmpB, sbits, swcnt, swscan, tileflags, tilemax;
    DevBuf &lab = F.lab, &lv = F.lv, &tile_seed = F.tile_seed, &totals = F.totals, &fstats = F.fstats;
    BS_TRY(msk.alloc(P_pix, s));
    BS_TRY(g.alloc(P_pix * 2, s));
    BS_TRY(d2.alloc(P_pix * 4, s));
    BS_TRY(tmpA.alloc(P_pix * 4, s));
    BS_TRY(lab.alloc(P_pix * 4, s));
    BS_TRY(lv.alloc(P_pix * 4, s));
    const size_t nsw = ((size_t)P_pix + 31) / 32 + 1;   // seed bitmap (one bit per tile pixel)
    BS_TRY(sbits.alloc_zero(4 * nsw, s));
    BS_TRY(swcnt.alloc(4 * nsw, s));
    BS_TRY(swscan.alloc(4 * nsw, s));
    BS_TRY(tileflags.alloc_zero(4 * ntiles, s));
    BS_TRY(tilemax.alloc_zero(4 * (ntiles + 1), s));

    // ---- mask, exact squared EDT
    g_prof.mark("s1.mask_rowdist", s);
    {
        dim3 gr((unsigned)rowdist_grid_rows(tiles), ntiles);
        ShiftView S;
        memset(&S, 0, sizeof(S));
        S.has_bias = cfg.has_bias, S.has_eps = cfg.has_seed_eps;
        for (int d = 0; d < 3; d++) S.bias[d] = cfg.bias[d];
        S.eps = cfg.seed_eps;
        S.has_sigma = cfg.has_sigma;
        DevBuf D2, d_pre, gA, gB, gW;
        PreTiles R;
        S.has_noise = cfg.has_noise, S.noise_eps = cfg.noise_eps, S.noise_seed = cfg.noise_seed;
        if (cfg.has_seed_eps || cfg.has_sigma || cfg.has_noise) {
            g_prof.mark("s1.shift_prepass", s);
            BS_TRY(build_pretiles(P, bidx, R, s));
            if (cfg.has_seed_eps) {
                BS_TRY(seed_distance_prepass<T>(P, R, A, D2, s));
                S.D2 = D2.as<uint32_t>();
            }
            if (cfg.has_sigma) {
                BS_TRY(gauss_prepass<T>(P, R, A, gA, gB, gW, &S.G, s));
                S.gstride = (size_t)R.Ppre;
            }
            std::vector<PreRef> trefs(ntiles);
            for (int i = 0; i < ntiles; i++) trefs[i] = R.refs[tiles[i].block];
            BS_TRY(d_pre.alloc(sizeof(PreRef) * ntiles, s));
            BS_CUDA(cudaMemcpyAsync(d_pre.p, trefs.data(), sizeof(PreRef) * ntiles, cudaMemcpyHostToDevice, s));
            BS_CUDA(cudaStreamSynchronize(s));   // trefs is a host-staged copy
            S.pre = d_pre.as<PreRef>();
            g_prof.mark("s1.mask_rowdist", s);
        }
        if (cfg.has_bias || cfg.has_seed_eps || cfg.has_sigma || cfg.has_noise)
            BS_LAUNCH((k_mask_rowdist<T, 1>), gr, 256, 0, s, dt, A, S, nullptr, msk.as<uint8_t>(), g.as<uint16_t>(),
                      tileflags.as<uint32_t>());
        else
            BS_LAUNCH((k_mask_rowdist<T, 0>), gr, 256, 0, s, dt, A, S, nullptr, msk.as<uint8_t>(), g.as<uint16_t>(),
                      tileflags.as<uint32_t>());
    }
    g_prof.mark("s1.edt", s);
    TileDims td;
    td.ntiles = ntiles, td.maxpix = maxpix, td.maxD = td.maxH = td.maxW = 0;
    for (auto &t : tiles) td.maxH = std::max(td.maxH, t.H), td.maxW = std::max(td.maxW, t.W), td.maxD = std::max(td.maxD, t.D);
    BS_TRY(launch_edt(dt, td, !xy, g.as<uint16_t>(), tmpA.as<uint32_t>(), d2.as<uint32_t>(), tilemax.as<uint32_t>(), s));
    // ---- maximum filter -> seeds (parent array in lv, seed bitmap in sbits)
    g_prof.mark("s1.maxfilt", s);
    if (!xy) BS_TRY(tmpB.alloc(P_pix * 4, s));
    BS_TRY(launch_seeds(dt, td, !xy, cfg.min_seed_distance, d2.as<uint32_t>(), msk.as<uint8_t>(), tmpA.as<uint32_t>(),
                        tmpB.as<uint32_t>(), lv.as<uint32_t>(), sbits.as<uint32_t>(), s));
    tmpB.release();
    g_prof.mark("s1.seed_cc", s);
    BS_LAUNCH(k_seed_union, grid, 256, 0, s, dt, lv.as<uint32_t>());

    // ---- histogram sizes (host sync #1: total histogram entries, number of seed pixels)
    DevBuf hsize, hbase;
    BS_TRY(hsize.alloc(4 * (ntiles + 1), s));
    BS_TRY(hbase.alloc(4 * (ntiles + 1), s));
    BS_TRY(totals.alloc_zero(4 * 8, s));
    uint32_t *d_tot = totals.as<uint32_t>();  // [0]=hist entries [1]=seed pixels [2]=levels [3]=mask pixels [4]=roots
    BS_LAUNCH(k_tile_hsize, cdiv(ntiles, 256), 256, 0, s, tilemax.as<uint32_t>(), hsize.as<uint32_t>(), ntiles);
    BS_TRY(scan_exclusive_u32(hsize.as<uint32_t>(), hbase.as<uint32_t>(), ntiles, d_tot + 0, s));
    BS_LAUNCH(k_popc_words, cdiv(nsw, 256), 256, 0, s, sbits.as<uint32_t>(), swcnt.as<uint32_t>(), nsw);
    BS_TRY(scan_exclusive_u32(swcnt.as<uint32_t>(), swscan.as<uint32_t>(), nsw, d_tot + 1, s));
    BS_LAUNCH(k_tile_seed_max, cdiv(ntiles, 256), 256, 0, s, dt, ntiles, sbits.as<uint32_t>(), swscan.as<uint32_t>(), d_tot + 1,
              tilemax.as<uint32_t>(), d_tot + 5, d_tot + 6);
    uint32_t h_tot[8];
    BS_CUDA(cudaMemcpyAsync(h_tot, d_tot, 32, cudaMemcpyDeviceToHost, s));
    BS_CUDA(cudaStreamSynchronize(s));
    const size_t Htot = h_tot[0], nseeds = h_tot[1];
    F.nseeds = nseeds;
    // flood v2 (on-chip state) needs 2-D tiles of <= 2^17 pixels, 15-bit labels and 16-bit levels
    const bool v2 = xy && g_flood_version != 1 && g_flood_version != 6 && maxpix <= F2_MAXPIX && h_tot[5] < 32767 && h_tot[6] < 65535;
    F.v2 = v2;

    g_prof.mark("s1.levels", s);
    DevBuf hist, nz, lrank, qoff, lvl_qstart, lvl_head, lvl_tail, tile_lvl, seedlist, queue, tile_q, lv16, availw;
    BS_TRY(tile_q.alloc(4 * (ntiles + 1), s));
    if (v2) {
        BS_TRY(lv16.alloc(2 * (size_t)P_pix, s));
        BS_TRY(availw.alloc_zero(4 * ((size_t)P_pix / 32 + 1), s));
    }
    BS_TRY(hist.alloc_zero(4 * (Htot + 1), s));
    BS_TRY(nz.alloc(Htot + 1, s));
    BS_TRY(lrank.alloc(4 * (Htot + 1), s));
    BS_TRY(qoff.alloc(4 * (Htot + 1), s));
    BS_TRY(lvl_qstart.alloc(4 * (Htot + 1), s));
    BS_TRY(lvl_head.alloc_zero(4 * (Htot + 1), s));
    BS_TRY(lvl_tail.alloc_zero(4 * (Htot + 1), s));
    BS_TRY(tile_lvl.alloc(4 * (ntiles + 1), s));
    BS_TRY(tile_seed.alloc(4 * (ntiles + 1), s));
    BS_TRY(seedlist.alloc(4 * (nseeds + 1), s));
    if (v2)
        BS_TRY(queue.alloc_fill(4 * (size_t)P_pix, 0xFF, s));
    else
        BS_TRY(queue.alloc(4 * (size_t)P_pix, s));
    BS_TRY(fstats.alloc_zero(16, s));
    // flood v2 labels of a tile fit shared memory as 16-bit values: the scatter after the flood rewrites the whole plane
    const size_t scat_smem = (((size_t)maxpix + 1) / 2) * 4;
    const bool tile_scatter = v2 && scat_smem <= 220 * 1024;
    BS_LAUNCH(k_seed_label_hist, grid, 256, 0, s, dt, lv.as<uint32_t>(), msk.as<uint8_t>(), d2.as<uint32_t>(),
              lab.as<uint32_t>(), hbase.as<uint32_t>(), hist.as<uint32_t>(), sbits.as<uint32_t>(), swscan.as<uint32_t>(),
              v2 ? (tile_scatter && !g_debug ? 2 : 1) : 0);
    if (Htot) {
        BS_LAUNCH(k_nzflag, cdiv(Htot, 256), 256, 0, s, hist.as<uint32_t>(), nz.as<uint8_t>(), Htot);
        BS_TRY(scan_exclusive_u8(nz.as<uint8_t>(), lrank.as<uint32_t>(), Htot, d_tot + 2, s));
        BS_TRY(scan_exclusive_u32(hist.as<uint32_t>(), qoff.as<uint32_t>(), Htot, d_tot + 3, s));
        BS_LAUNCH(k_levels, cdiv(Htot, 256), 256, 0, s, hist.as<uint32_t>(), lrank.as<uint32_t>(), qoff.as<uint32_t>(),
                  lvl_qstart.as<uint32_t>(), Htot);
    }
    // lrank needs a valid entry at hbase[t] for every tile: hbase[t] < Htot always (hsize >= 1)
    BS_LAUNCH(k_tile_ranges, cdiv(ntiles + 1, 256), 256, 0, s, dt, ntiles, hbase.as<uint32_t>(), lrank.as<uint32_t>(),
              d_tot + 2, sbits.as<uint32_t>(), swscan.as<uint32_t>(), d_tot + 1, tile_lvl.as<uint32_t>(), tile_seed.as<uint32_t>(),
              qoff.as<uint32_t>(), d_tot + 3, tile_q.as<uint32_t>());
    // the seed parent array lives in lv and is consumed by k_seed_label_hist above; now lv becomes the level
    BS_LAUNCH(k_pixel_levels, grid, 256, 0, s, dt, msk.as<uint8_t>(), d2.as<uint32_t>(), hbase.as<uint32_t>(),
              lrank.as<uint32_t>(), lv.as<uint32_t>(), sbits.as<uint32_t>(), swscan.as<uint32_t>(), seedlist.as<uint32_t>(),
              v2 ? lv16.as<uint16_t>() : nullptr, v2 ? availw.as<uint32_t>() : nullptr);
    if (g_debug) {
        DevBuf c1, c2;
        BS_TRY(c1.alloc(P_pix * 4, s));
        BS_CUDA(cudaMemcpyAsync(c1.p, d2.p, P_pix * 4, cudaMemcpyDeviceToDevice, s));
        keep_debug(P, "d2", c1, 4, P_pix);
        BS_TRY(c2.alloc(P_pix * 4, s));
        BS_CUDA(cudaMemcpyAsync(c2.p, lab.p, P_pix * 4, cudaMemcpyDeviceToDevice, s));
        keep_debug(P, "seeds", c2, 4, P_pix);
    }
    // ---- flood
    g_prof.mark("s1.flood", s);
    // tiles flood v2 cannot take go to the CTA-per-tile kernel when the tile's level tables fit in shared memory
    // (host sync: the largest number of levels in one tile), else to the one-warp kernel
    int levcap3 = 0;
    bool flood3 = false;
    if (!v2 && g_flood_version != 1 && g_flood_version != 6) {
        BS_LAUNCH(k_tile_nlev_max, cdiv(ntiles, 256), 256, 0, s, tile_lvl.as<uint32_t>(), ntiles, d_tot + 7);
        uint32_t h_nlev = 0;
        BS_CUDA(cudaMemcpyAsync(&h_nlev, d_tot + 7, 4, cudaMemcpyDeviceToHost, s));
        BS_CUDA(cudaStreamSynchronize(s));
        levcap3 = (int)h_nlev + 1;
        flood3 = flood3_smem(levcap3) <= 227 * 1024 && maxpix <= (long long)F3_PIXMASK;
    }
    if (v2) {
        const int nwords_max = (int)((maxpix + 31) / 32);
        // bitmap in shared memory only when every tile of the batch is resident then anyway (16 tiles per SM); otherwise
        // it stays in global memory and all tiles run in one wave
        const bool gavail = g_flood_version == 3 || g_flood_version == 4 || (g_flood_version != 2 && ntiles > 148 * 16);
        // level tails + occupancy bits in shared memory when they leave room for every tile of the batch to be resident
        // (host sync: the largest number of levels in one tile)
        int levcap = 0;
        bool slev = false;
        if (gavail && g_flood_version != 3) {
            BS_LAUNCH(k_tile_nlev_max, cdiv(ntiles, 256), 256, 0, s, tile_lvl.as<uint32_t>(), ntiles, d_tot + 7);
            uint32_t h_nlev = 0;
            BS_CUDA(cudaMemcpyAsync(&h_nlev, d_tot + 7, 4, cudaMemcpyDeviceToHost, s));
            BS_CUDA(cudaStreamSynchronize(s));
            levcap = (int)h_nlev + 1;
            const size_t per_tile = (size_t)(F2_HASH + levcap + (levcap + 31) / 32) * 4 + 1024;   // + 1 KB the driver reserves per CTA
            const size_t tiles_per_sm = (size_t)(ntiles + 147) / 148;
            slev = per_tile * std::min<size_t>(tiles_per_sm, 32) <= 227 * 1024;
        }
        const size_t smem = gavail ? (size_t)F2_WARPS * (F2_HASH + (slev ? levcap + (levcap + 31) / 32 : 0)) * 4
                                   : (size_t)F2_WARPS * (nwords_max + F2_HASH) * 4;
        static std::atomic<unsigned long long> attr_set(0);   // one bit per device: function attributes are per device
        if (!dev_once(attr_set)) {
            BS_CUDA(cudaFuncSetAttribute(k_flood2<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        }
#define BS_FLOOD2(G_, S_)                                                                                                    \
    BS_LAUNCH((k_flood2<G_, S_>), cdiv(ntiles, F2_WARPS), 32 * F2_WARPS, smem, s, dt, ntiles, lab.as<uint32_t>(),             \
              lv16.as<uint16_t>(), availw.as<uint32_t>(), queue.as<uint32_t>(), lvl_qstart.as<uint32_t>(),                   \
              lvl_head.as<uint32_t>(), lvl_tail.as<uint32_t>(), tile_lvl.as<uint32_t>(), seedlist.as<uint32_t>(),            \
              tile_seed.as<uint32_t>(), nwords_max, levcap, fstats.as<uint32_t>())
        if (slev)
            BS_FLOOD2(true, true);
        else if (gavail)
            BS_FLOOD2(true, false);
        else
            BS_FLOOD2(false, false);
#undef BS_FLOOD2
        g_prof.mark("s1.flood_scatter", s);
        if (tile_scatter) {
            BS_CUDA(cudaFuncSetAttribute(k_scatter_labels_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scat_smem));
            BS_LAUNCH(k_scatter_labels_tile, ntiles, SCAT_NT, scat_smem, s, dt, tile_q.as<uint32_t>(), queue.as<uint32_t>(), lab.as<uint32_t>());
        } else
        BS_LAUNCH(k_scatter_labels, dim3((unsigned)std::min<long long>(std::max<long long>((maxpix + 1023) / 1024, 1), 2048), ntiles), 256, 0, s, dt, tile_q.as<uint32_t>(), queue.as<uint32_t>(), lab.as<uint32_t>());
    } else if (flood3) {
        BS_CUDA(cudaFuncSetAttribute(k_flood3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)flood3_smem(levcap3)));
        BS_LAUNCH(k_flood3, ntiles, F3_NT, flood3_smem(levcap3), s, dt, ntiles, lab.as<uint32_t>(), lv.as<uint32_t>(),
                  queue.as<uint32_t>(), lvl_qstart.as<uint32_t>(), tile_lvl.as<uint32_t>(), seedlist.as<uint32_t>(),
                  tile_seed.as<uint32_t>(), levcap3, fstats.as<uint32_t>());
    } else if (g_flood_version == 6) {
        // the faithful heap (skimage's own seed ties): three 32-bit planes of heap storage, a tile's heap at its pixel base
        DevBuf hl, ha, hp;
        BS_TRY(hl.alloc(4 * (size_t)P_pix, s));
        BS_TRY(ha.alloc(4 * (size_t)P_pix, s));
        BS_TRY(hp.alloc(4 * (size_t)P_pix, s));
        BS_LAUNCH(k_flood_heap, cdiv((size_t)ntiles, 32), 32, 0, s, dt, ntiles, lab.as<uint32_t>(), lv.as<uint32_t>(), seedlist.as<uint32_t>(),
                  tile_seed.as<uint32_t>(), hl.as<uint32_t>(), ha.as<uint32_t>(), hp.as<uint32_t>());
        BS_CUDA(cudaStreamSynchronize(s));
    } else {
        BS_LAUNCH(k_flood, cdiv((size_t)ntiles * 32, 64), 64, 0, s, dt, ntiles, lab.as<uint32_t>(), lv.as<uint32_t>(),
                  queue.as<uint32_t>(), lvl_qstart.as<uint32_t>(), lvl_head.as<uint32_t>(), lvl_tail.as<uint32_t>(),
                  tile_lvl.as<uint32_t>(), seedlist.as<uint32_t>(), tile_seed.as<uint32_t>(), fstats.as<uint32_t>());
    }
    // release what the flood no longer needs
    queue.release();
    hist.release();
    nz.release();
    lrank.release();
    qoff.release();
    g.release();
    tmpA.release();
    if (g_debug) {
        DevBuf c1;
        BS_TRY(c1.alloc(P_pix * 4, s));
        BS_CUDA(cudaMemcpyAsync(c1.p, lab.p, P_pix * 4, cudaMemcpyDeviceToDevice, s));
        keep_debug(P, "flood", c1, 4, P_pix);
        keep_debug(P, "flood_stats", fstats, 4, 4);
    }
    return BS_OK;
}

extern int g_front_version;   // 0 = automatic, 1 = unfused chain only, 2 = fused without TMA, 3 = fused, TMA required when eligible

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*tmap_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tmap_encode_fn get_tmap_encode() {
    static tmap_encode_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (tmap_encode_fn)sym;
        else
            cudaGetLastError();
    }
    return fn;
}

// The fused front end (front2d.cuh) for batches of 2-D tiles of unshifted affinities whose 16-bit planes fit one CTA's
// shared memory.  *done = false: not eligible, or a tile overflowed one of the on-chip tables (the caller then runs the
// unfused chain on the same batch).
template <typename T>
static int stage1_front_fused(Plan &P, AffView A, const std::vector<Tile> &tiles, const Tile *dt, long long P_pix, long long maxpix,
                              S1Front &F, bool *done, cudaStream_t s) {
    const bs_ws_config &cfg = P.cfg;
    *done = false;
    const int ntiles = (int)tiles.size();
    if (!cfg.fragments_in_xy || cfg.has_bias || cfg.has_seed_eps || cfg.has_sigma || cfg.has_noise || g_debug || g_front_version == 1 ||
        g_flood_version != 0 || maxpix > F2_MAXPIX)
        return BS_OK;
    int maxH = 0, maxW = 0;
    for (auto &t : tiles) maxH = std::max(maxH, t.H), maxW = std::max(maxW, t.W);
    if (maxH > 128 * FR_MAXG || maxW > 1024 || cfg.min_seed_distance > 64) return BS_OK;
    // shared memory of one tile CTA: 16-bit plane + bitmap + scratch (seed ranks / union-find, then level tables)
    size_t need_fixed = 0, nwords_max = 0;
    for (auto &t : tiles) {
        const size_t nw = (size_t)t.H * ((t.W + 31) / 32);
        need_fixed = std::max(need_fixed, (((size_t)t.H * fr_pitch(t.W) * 2 + 15) & ~(size_t)15) + nw * 4);
        nwords_max = std::max(nwords_max, nw);
    }
    if (need_fixed + 3072 + 4 * 512 > (size_t)FR_SMEM_TOTAL) return BS_OK;
    const int scr_bytes = (int)(((size_t)FR_SMEM_TOTAL - need_fixed) & ~(size_t)15);
    if ((size_t)scr_bytes < 3072 + (((size_t)maxH * 2 + 15) & ~(size_t)15) + 4 * 512) return BS_OK;
    if ((scr_bytes - 3072) / (2 * fr_pitch(maxW)) < cfg.min_seed_distance - 1 + 4) return BS_OK;   // rows of the maximum filter's band buffer
    const int levtab = std::min(4096, (scr_bytes - 3072) / 4);   // the kernel's level-count table (scr_work / 4 entries)

    // ---- mask bitmap
    g_prof.mark("s1.mask_bits", s);
    std::vector<uint32_t> h_mbase(ntiles + 1);
    size_t mwords = 0;
    for (int i = 0; i < ntiles; i++) {
        h_mbase[i] = (uint32_t)mwords;
        mwords += (size_t)tiles[i].H * ((tiles[i].W + 31) / 32);
    }
    h_mbase[ntiles] = (uint32_t)mwords;
    BS_ARG(mwords < (1ull << 32), "stage1: mask bitmap exceeds 32-bit indexing");
    DevBuf d_amap, mbase, mbits, lv16, availw, seedent, queue, lvl_qstart, lvl_head, t_nseeds, t_nlev, t_nmask, flags;
    BS_TRY(mbase.alloc(4 * (size_t)(ntiles + 1), s));
    BS_CUDA(cudaMemcpyAsync(mbase.p, h_mbase.data(), 4 * (size_t)(ntiles + 1), cudaMemcpyHostToDevice, s));
    BS_TRY(mbits.alloc(4 * mwords, s));
    BS_TRY(lv16.alloc(2 * (size_t)P_pix + 64, s));
    BS_TRY(availw.alloc(4 * ((size_t)P_pix / 32 + 2), s));
    BS_TRY(seedent.alloc(4 * (size_t)P_pix, s));
    BS_TRY(queue.alloc_fill(4 * (size_t)P_pix, 0xFF, s));
    BS_TRY(lvl_qstart.alloc(4 * (size_t)ntiles * levtab, s));
    BS_TRY(lvl_head.alloc(4 * (size_t)ntiles * levtab, s));
    BS_TRY(t_nseeds.alloc_zero(4 * (size_t)(ntiles + 1), s));
    BS_TRY(t_nlev.alloc(4 * (size_t)(ntiles + 1), s));
    BS_TRY(t_nmask.alloc(4 * (size_t)(ntiles + 1), s));
    BS_TRY(flags.alloc_zero(32, s));
    BS_TRY(F.totals.alloc_zero(32, s));
    BS_TRY(F.fstats.alloc_zero(16, s));
    BS_TRY(F.tile_seed.alloc(4 * (size_t)(ntiles + 1), s));
    const uint32_t *d_mbase = mbase.as<uint32_t>();
    bool used_tma = false;
    // TMA boxes: rows of uint32 elements, pitch = the tile width + 15 bytes of misalignment + 16 bytes for the consumer's
    // second aligned read, rounded up to 16 bytes (box rows are dense in shared memory)
    const int tma_pitch = ((maxW + 31) + 15) & ~15;
    if (sizeof(T) == 1 && g_front_version != 2 && !A.mask && A.X % 16 == 0 && ((uintptr_t)A.p & 15) == 0 && tma_pitch <= 1024 &&
        (size_t)A.Zw * A.Y * A.X < (1ull << 32)) {
        tmap_encode_fn enc = get_tmap_encode();
        if (enc) {
            CUtensorMap amap;
            const cuuint64_t gdim[4] = {(cuuint64_t)A.X / 4, (cuuint64_t)A.Y, (cuuint64_t)A.Zw, (cuuint64_t)A.C};
            const cuuint64_t gstr[3] = {(cuuint64_t)A.X, (cuuint64_t)A.X * A.Y, (cuuint64_t)A.X * A.Y * A.Zw};
            const cuuint32_t box[4] = {(cuuint32_t)(tma_pitch / 4), (cuuint32_t)TB_ROWS, 1, 1};
            const cuuint32_t estr[4] = {1, 1, 1, 1};
            CUresult r = enc(&amap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, const_cast<void *>(A.p), gdim, gstr, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r == CUDA_SUCCESS) {
                // the descriptor lives in global memory (64-byte aligned), the kernel gets its address
                BS_TRY(d_amap.alloc(sizeof(CUtensorMap) + 64, s));
                CUtensorMap *dmap = (CUtensorMap *)(((uintptr_t)d_amap.p + 63) & ~(uintptr_t)63);
                BS_CUDA(cudaMemcpyAsync(dmap, &amap, sizeof(CUtensorMap), cudaMemcpyHostToDevice, s));
                BS_CUDA(cudaStreamSynchronize(s));   // amap is a host-staged copy
                const int strips = (maxH + TB_ROWS - 1) / TB_ROWS;
                const size_t smem = (size_t)TB_STAGES * 2 * TB_ROWS * tma_pitch;
                const int nbricks = ntiles * strips;
                const int grid_tma = std::min(nbricks, 148 * 4);
                BS_CUDA(cudaFuncSetAttribute(k_mask_bits_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                BS_LAUNCH(k_mask_bits_tma, grid_tma, 256, smem, s, dmap, dt, d_mbase, ntiles, strips, A.z0, tma_pitch, mbits.as<uint32_t>());
                used_tma = true;
            }
        }
        BS_ARG(used_tma || g_front_version != 3, "stage1: the TMA mask kernel was required (front version 3) but the tensor map could not be encoded");
    }
    if (!used_tma) {
        const dim3 gr((unsigned)std::min(std::max((maxH + 7) / 8, 1), 64), ntiles);
        const dim3 gr4((unsigned)std::min(std::max((maxH + 31) / 32, 1), 16), ntiles);
        if (sizeof(T) == 1)
            BS_LAUNCH(k_mask_bits_u8, gr4, 256, 0, s, dt, d_mbase, A, mbits.as<uint32_t>());
        else
            BS_LAUNCH((k_mask_bits_generic<T>), gr, 256, 0, s, dt, d_mbase, A, mbits.as<uint32_t>());
    }
    // ---- the tile front end
    g_prof.mark("s1.tile_front", s);
    FrontOut O;
    O.lv16 = lv16.as<uint16_t>();
    O.availw = availw.as<uint32_t>();
    O.seedent = seedent.as<uint32_t>();
    O.lvl_qstart = lvl_qstart.as<uint32_t>();
    O.lvl_head = lvl_head.as<uint32_t>();
    O.t_nseeds = t_nseeds.as<uint32_t>();
    O.t_nlev = t_nlev.as<uint32_t>();
    O.t_nmask = t_nmask.as<uint32_t>();
    O.flags = flags.as<uint32_t>();
    O.levtab = levtab;
    const size_t fr_smem = need_fixed + (size_t)scr_bytes;
    BS_CUDA(cudaFuncSetAttribute(k_tile_front, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fr_smem));
    BS_LAUNCH(k_tile_front, ntiles, FR_NT, fr_smem, s, dt, d_mbase, mbits.as<uint32_t>(), cfg.min_seed_distance, scr_bytes, O);
    // fragment-table base per tile = seed pixels before the tile
    uint32_t *d_tot = F.totals.as<uint32_t>();
    BS_TRY(scan_exclusive_u32(t_nseeds.as<uint32_t>(), F.tile_seed.as<uint32_t>(), (size_t)ntiles + 1, d_tot + 1, s));
    // host sync: overflow flags, total seed pixels, the largest level count of a tile
    uint32_t h_flags[8], h_tot[8];
    BS_CUDA(cudaMemcpyAsync(h_flags, flags.p, 32, cudaMemcpyDeviceToHost, s));
    BS_CUDA(cudaMemcpyAsync(h_tot, d_tot, 32, cudaMemcpyDeviceToHost, s));
    BS_CUDA(cudaStreamSynchronize(s));
    if (h_flags[0] != 0) {
        // a tile did not fit the on-chip tables (squared distance >= 16384, too many seed pixels or levels): unfused chain
        F.tile_seed.release();
        F.totals.release();
        F.fstats.release();
        return BS_OK;
    }
    F.nseeds = h_tot[1];
    F.v2 = true;
    // ---- flood
    g_prof.mark("s1.flood", s);
    {
        const int levcap = (int)h_flags[1] + 1;
        const size_t smem = (size_t)(F2_HASH + levcap + (levcap + 31) / 32) * 4;
        BS_LAUNCH((k_flood2<true, true, true>), ntiles, 32, smem, s, dt, ntiles, nullptr, lv16.as<uint16_t>(), availw.as<uint32_t>(),
                  queue.as<uint32_t>(), lvl_qstart.as<uint32_t>(), lvl_head.as<uint32_t>(), nullptr, nullptr, nullptr, nullptr, 0, levcap,
                  F.fstats.as<uint32_t>(), seedent.as<uint32_t>(), t_nlev.as<uint32_t>(), t_nseeds.as<uint32_t>(), levtab);
    }
    g_prof.mark("s1.flood_scatter", s);
    BS_TRY(F.lab.alloc(2 * (size_t)P_pix + 64, s));
    BS_TRY(F.lv.alloc(4 * (size_t)P_pix, s));
    F.lab16 = true;
    {
        const size_t scat_smem = (((size_t)maxpix + 1) / 2) * 4;
        BS_CUDA(cudaFuncSetAttribute(k_scatter_labels_tile5, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scat_smem));
        BS_LAUNCH(k_scatter_labels_tile5, ntiles, SCAT_NT, scat_smem, s, dt, t_nmask.as<uint32_t>(), queue.as<uint32_t>(),
                  F.lab.as<uint16_t>());
    }
    *done = true;
    return BS_OK;
}

template <typename T>
static int stage1_batch(Plan &P, const std::vector<int> &bidx, AffView A, uint64_t *frags_out, long long node_base,
                        long long *n_new_nodes, cudaStream_t s, const uint32_t *ext_labels = nullptr, size_t ext_nlabels = 0) {
    typedef typename AffOps<T>::acc_t acc_t;
    const bs_ws_config &cfg = P.cfg;
    // ext_labels: the fragments of every block's read ROI are given (bs_stage1_from_labels: one (rz, ry, rx) u32 volume
    // per block of the batch, packed back to back, values 0 or 1..ext_nlabels unique over the batch) -- the blocks are
    // then 3-D tiles whatever fragments_in_xy says: the given fragments may span z slices
    const bool xy = cfg.fragments_in_xy != 0 && !ext_labels;
    // ---- tiles
    std::vector<Tile> tiles;
    std::vector<BlkDev> blks;
    long long P_pix = 0, V_w = 0, maxpix = 0, maxw = 0;
    for (size_t bi = 0; bi < bidx.size(); bi++) {
        const Blk &b = P.blocks[bidx[bi]];
        BlkDev bd;
        bd.block_id = b.block_id;
        bd.wbase = V_w;
        for (int d = 0; d < 3; d++) bd.wo[d] = b.wo[d], bd.ws[d] = b.ws[d];
        bd.plan_index = bidx[bi];
        bd.pad_ = 0;
        blks.push_back(bd);
        long long wv = (long long)b.ws[0] * b.ws[1] * b.ws[2];
        maxw = std::max(maxw, wv);
        if (xy) {
            for (int z = 0; z < b.ws[0]; z++) {
                Tile t;
                t.gz = b.wo[0] + z, t.gy = b.ro[1], t.gx = b.ro[2];
                t.D = 1, t.H = b.rs[1], t.W = b.rs[2];
                t.wz = 0, t.wy = cfg.context[1], t.wx = cfg.context[2];
                t.wD = 1, t.wH = b.ws[1], t.wW = b.ws[2];
                t.block = (int)bi;
                t.ndim = 2;
                t.base = P_pix;
                t.wbase = V_w + (long long)z * b.ws[1] * b.ws[2];
                long long np = (long long)t.H * t.W;
                P_pix += (np + 31) & ~31LL;   // 32-aligned tile bases (bitmap words of flood v2)
                maxpix = std::max(maxpix, np);
                t.set_divs();
                tiles.push_back(t);
            }
        } else {
            Tile t;
            t.gz = b.ro[0], t.gy = b.ro[1], t.gx = b.ro[2];
            t.D = b.rs[0], t.H = b.rs[1], t.W = b.rs[2];
            t.wz = cfg.context[0], t.wy = cfg.context[1], t.wx = cfg.context[2];
            t.wD = b.ws[0], t.wH = b.ws[1], t.wW = b.ws[2];
            t.block = (int)bi;
            t.ndim = 3;
            t.base = P_pix;
            t.wbase = V_w;
            long long np = (long long)t.D * t.H * t.W;
            P_pix += (np + 31) & ~31LL;   // 32-aligned tile bases (16-byte label loads of flood v3)
            maxpix = std::max(maxpix, np);
            t.set_divs();
                tiles.push_back(t);
        }
        V_w += wv;
    }
    const int ntiles = (int)tiles.size();
    BS_ARG(ntiles > 0 && ntiles <= 65535, "stage1: batch has too many tiles (lower max_batch_voxels)");
    BS_ARG(P_pix < (1LL << 31) && maxpix < (1LL << 31), "stage1: batch too large for 32-bit tile indices");
    for (auto &t : tiles) BS_ARG(t.W <= MAXW && t.W < 65535, "stage1: tile wider than 4096 voxels is not supported");

    DevBuf d_tiles, d_blks;
    BS_TRY(d_tiles.alloc(sizeof(Tile) * ntiles, s));
    BS_TRY(d_blks.alloc(sizeof(BlkDev) * blks.size(), s));
    BS_CUDA(cudaMemcpyAsync(d_tiles.p, tiles.data(), sizeof(Tile) * ntiles, cudaMemcpyHostToDevice, s));
    BS_CUDA(cudaMemcpyAsync(d_blks.p, blks.data(), sizeof(BlkDev) * blks.size(), cudaMemcpyHostToDevice, s));
    const Tile *dt = d_tiles.as<Tile>();

    const unsigned gx = (unsigned)std::min<long long>(std::max<long long>((maxpix + PIX_PER_CTA - 1) / PIX_PER_CTA, 1), 2048);
    const dim3 grid(gx, ntiles);

    S1Front F;
    bool front_done = false;
    if (ext_labels) {
        BS_TRY(F.lab.alloc(4 * (size_t)P_pix, s));
        BS_TRY(F.lv.alloc(4 * (size_t)P_pix, s));
        BS_TRY(F.tile_seed.alloc_zero(4 * (size_t)(ntiles + 1), s));   // one fragment table for the batch: index = label - 1
        BS_TRY(F.totals.alloc_zero(32, s));
        BS_TRY(F.fstats.alloc_zero(16, s));
        size_t src = 0;
        for (auto &t : tiles) {
            const size_t np = (size_t)t.D * t.H * t.W;
            BS_CUDA(cudaMemcpyAsync(F.lab.as<uint32_t>() + t.base, ext_labels + src, 4 * np, cudaMemcpyDeviceToDevice, s));
            src += np;
        }
        F.nseeds = ext_nlabels;
        F.v2 = true;
        front_done = true;
    }
    if (!front_done) BS_TRY((stage1_front_fused<T>(P, A, tiles, dt, P_pix, maxpix, F, &front_done, s)));
    if (!front_done) BS_TRY((stage1_front_unfused<T>(P, bidx, A, tiles, dt, P_pix, maxpix, grid, F, s)));
    DevBuf &lab = F.lab, &lv = F.lv, &tile_seed = F.tile_seed, &totals = F.totals;
    uint32_t *d_tot = totals.as<uint32_t>();
    uint32_t h_tot[8];
    const size_t nseeds = F.nseeds;
    const bool v2 = F.v2;


    // ---- fragment statistics, keep/drop, crop relabel
    g_prof.mark("s1.fragstats", s);
    DevBuf fsum, fcnt, fmin, fflag, fbase_v1, bits, wcnt, wscan;
    const bool need_stats = cfg.filter_fragments > 0.0 || cfg.remove_debris > 0;
    // fragment table: flood v2 labels are ranks among the tile's seed pixels, v1 labels are root pixel + 1
    const size_t nF = (v2 ? nseeds : (size_t)P_pix) + 1;
    const uint32_t *d_fbase = tile_seed.as<uint32_t>();
    if (!v2) {
        std::vector<uint32_t> hb(ntiles);
        for (int i = 0; i < ntiles; i++) hb[i] = (uint32_t)tiles[i].base;
        BS_TRY(fbase_v1.alloc(4 * (size_t)ntiles, s));
        BS_CUDA(cudaMemcpyAsync(fbase_v1.p, hb.data(), 4 * (size_t)ntiles, cudaMemcpyHostToDevice, s));
        BS_CUDA(cudaStreamSynchronize(s));
        d_fbase = fbase_v1.as<uint32_t>();
    }
    BS_TRY(fsum.alloc_zero(nF * 8, s));
    BS_TRY(fcnt.alloc_zero(nF * 4, s));
    BS_TRY(fmin.alloc_fill(nF * 4, 0xFF, s));
    BS_TRY(fflag.alloc_zero(nF, s));
    const bool lab16 = F.lab16;
    if (lab16)
        BS_LAUNCH((k_fragstats<T, uint16_t>), grid, 256, 0, s, dt, A, lab.as<uint16_t>(), d_fbase, need_stats ? 1 : 0, ext_labels ? 1 : 0, fsum.as<acc_t>(),
                  fcnt.as<uint32_t>(), fmin.as<uint32_t>(), fflag.as<uint8_t>());
    else
        BS_LAUNCH((k_fragstats<T, uint32_t>), grid, 256, 0, s, dt, A, lab.as<uint32_t>(), d_fbase, need_stats ? 1 : 0, ext_labels ? 1 : 0, fsum.as<acc_t>(),
                  fcnt.as<uint32_t>(), fmin.as<uint32_t>(), fflag.as<uint8_t>());
    BS_LAUNCH((k_frag_decide<acc_t>), cdiv(nF, 256), 256, 0, s, fsum.as<acc_t>(), fcnt.as<uint32_t>(), fflag.as<uint8_t>(), nF,
              need_stats ? cfg.filter_fragments : 0.0, need_stats ? cfg.remove_debris : 0, sizeof(T) == 1 ? 1 : 0);
    g_prof.mark("s1.crop_cc", s);
    long long maxwt = 1;
    for (auto &t : tiles) maxwt = std::max(maxwt, (long long)t.wD * t.wH * t.wW);
    // the union-find kernels run better with many small CTAs (their atomics stay spatially clustered)
    const dim3 gridw((unsigned)std::min<long long>(std::max<long long>((maxwt + 1023) / 1024, 1), 2048), ntiles);
    const dim3 gridf((unsigned)std::min<long long>(std::max<long long>((maxwt + PIX_PER_CTA - 1) / PIX_PER_CTA, 1), 2048), ntiles);
    // cpar reuses lv
    const size_t nwords = ((size_t)V_w + 31) / 32 + 1;
    DevBuf xbits;
    BS_TRY(xbits.alloc_zero(4 * (nwords + 1), s));
    if (lab16) {
        BS_LAUNCH(k_crop_init<uint16_t>, gridw, 256, 0, s, dt, lab.as<uint16_t>(), d_fbase, fflag.as<uint8_t>(), lv.as<uint32_t>(), xbits.as<uint32_t>());
        BS_LAUNCH(k_crop_union<uint16_t>, gridw, 256, 0, s, dt, lab.as<uint16_t>(), lv.as<uint32_t>(), xbits.as<uint32_t>());
    } else {
        BS_LAUNCH(k_crop_init<uint32_t>, gridw, 256, 0, s, dt, lab.as<uint32_t>(), d_fbase, fflag.as<uint8_t>(), lv.as<uint32_t>(), xbits.as<uint32_t>());
        BS_LAUNCH(k_crop_union<uint32_t>, gridw, 256, 0, s, dt, lab.as<uint32_t>(), lv.as<uint32_t>(), xbits.as<uint32_t>());
    }
    BS_TRY(bits.alloc_zero(4 * nwords, s));
    BS_TRY(wcnt.alloc(4 * nwords, s));
    BS_TRY(wscan.alloc(4 * nwords, s));
    BS_LAUNCH(k_root_bits_pix, gridw, 256, 0, s, dt, lv.as<uint32_t>(), xbits.as<uint32_t>(), bits.as<uint32_t>());
    BS_LAUNCH(k_root_bits_frag, cdiv(nF, 256), 256, 0, s, fcnt.as<uint32_t>(), fflag.as<uint8_t>(), fmin.as<uint32_t>(), nF,
              bits.as<uint32_t>());
    BS_LAUNCH(k_popc_words, cdiv(nwords, 256), 256, 0, s, bits.as<uint32_t>(), wcnt.as<uint32_t>(), nwords);
    BS_TRY(scan_exclusive_u32(wcnt.as<uint32_t>(), wscan.as<uint32_t>(), nwords, d_tot + 4, s));
    // host sync #2: number of fragments in this batch
    BS_CUDA(cudaMemcpyAsync(h_tot, d_tot, 32, cudaMemcpyDeviceToHost, s));
    BS_CUDA(cudaStreamSynchronize(s));
    const uint32_t nn = h_tot[4];
    *n_new_nodes = nn;

    g_prof.mark("s1.finalize", s);
    DevBuf ncnt, nsum, blk_first;
    BS_TRY(ncnt.alloc_zero(4 * ((size_t)nn + 1), s));
    BS_TRY(nsum.alloc_zero(24 * ((size_t)nn + 1), s));
    BS_TRY(blk_first.alloc(4 * blks.size(), s));
    BS_LAUNCH(k_blk_first, cdiv(blks.size(), 256), 256, 0, s, d_blks.as<BlkDev>(), (int)blks.size(), bits.as<uint32_t>(),
              wscan.as<uint32_t>(), blk_first.as<uint32_t>());
#define BS_FINALIZE(LT_)                                                                                                              \
    BS_LAUNCH(k_finalize<LT_>, gridf, 256, 0, s, dt, d_blks.as<BlkDev>(), lab.as<LT_>(), lv.as<uint32_t>(), d_fbase,                   \
              fflag.as<uint8_t>(), fmin.as<uint32_t>(), bits.as<uint32_t>(), wscan.as<uint32_t>(), blk_first.as<uint32_t>(),           \
              P.nvox_block, cfg.roi_offset[0] + (cfg.win_z > 0 ? cfg.win_z0 : 0), cfg.roi_offset[1], cfg.roi_offset[2],                 \
              cfg.roi_shape[1], cfg.roi_shape[2], frags_out, ncnt.as<uint32_t>(), nsum.as<unsigned long long>())
    if (lab16)
        BS_FINALIZE(uint16_t);
    else
        BS_FINALIZE(uint32_t);
#undef BS_FINALIZE
    // ---- grow the plan's node table
    {
        DevBuf nid, npos, nsz;
        size_t tot = (size_t)(node_base + nn);
        BS_TRY(nid.alloc_persistent(8 * (tot + 1), s));
        BS_TRY(npos.alloc_persistent(12 * (tot + 1), s));
        BS_TRY(nsz.alloc_persistent(4 * (tot + 1), s));
        if (node_base) {
            BS_CUDA(cudaMemcpyAsync(nid.p, P.node_id.p, 8 * node_base, cudaMemcpyDeviceToDevice, s));
            BS_CUDA(cudaMemcpyAsync(npos.p, P.node_pos.p, 12 * node_base, cudaMemcpyDeviceToDevice, s));
            BS_CUDA(cudaMemcpyAsync(nsz.p, P.node_size.p, 4 * node_base, cudaMemcpyDeviceToDevice, s));
        }
        if (nn)
            BS_LAUNCH(k_nodes, cdiv(nn, 256), 256, 0, s, d_blks.as<BlkDev>(), (int)blks.size(), blk_first.as<uint32_t>(), nn,
                      ncnt.as<uint32_t>(), nsum.as<unsigned long long>(), P.nvox_block, nid.as<uint64_t>() + node_base,
                      npos.as<int32_t>() + 3 * node_base, nsz.as<uint32_t>() + node_base);
        P.node_id.swap(nid);
        P.node_pos.swap(npos);
        P.node_size.swap(nsz);
    }
    // per-block counts
    std::vector<uint32_t> h_first(blks.size());
    BS_CUDA(cudaMemcpyAsync(h_first.data(), blk_first.p, 4 * blks.size(), cudaMemcpyDeviceToHost, s));
    BS_CUDA(cudaStreamSynchronize(s));
    for (size_t i = 0; i < blks.size(); i++) {
        uint32_t nxt = i + 1 < blks.size() ? h_first[i + 1] : nn;
        P.block_count[blks[i].plan_index] = (long long)nxt - h_first[i];
    }
    BS_CUDA(cudaGetLastError());
    return BS_OK;
}

int stage1_run(Plan &P, const void *affs, const uint8_t *mask, uint64_t *frags_out, cudaStream_t s, const uint32_t *ext_labels,
               size_t ext_nlabels) {
    const bs_ws_config &cfg = P.cfg;
    AffView A;
    A.p = affs;
    A.mask = mask;
    A.C = cfg.n_channels;
    A.Z = cfg.vol_shape[0], A.Y = cfg.vol_shape[1], A.X = cfg.vol_shape[2];
    A.z0 = cfg.win_z > 0 ? cfg.win_z0 : 0;
    A.Zw = cfg.win_z > 0 ? cfg.win_z : cfg.vol_shape[0];
    long long cap = cfg.max_batch_voxels > 0 ? cfg.max_batch_voxels : (1LL << 30);
    cap = std::min(cap, (1LL << 31) - 1);
    for (int bi : P.owned) P.block_count[bi] = 0;
    P.counts_global = false;
    P.n_nodes = 0;
    g_prof.reset();
    // batches of owned blocks (ascending block id)
    size_t i = 0;
    long long node_base = 0;
    while (i < P.owned.size()) {
        std::vector<int> batch;
        long long pix = 0;
        int tiles = 0;
        while (i < P.owned.size()) {
            const Blk &b = P.blocks[P.owned[i]];
            const bool xy_tiles = cfg.fragments_in_xy && !ext_labels;
            long long bp = xy_tiles ? (long long)b.ws[0] * b.rs[1] * b.rs[2] : (long long)b.rs[0] * b.rs[1] * b.rs[2];
            int bt = xy_tiles ? b.ws[0] : 1;
            if (!batch.empty() && (pix + bp > cap || tiles + bt > 65535)) break;
            batch.push_back(P.owned[i]);
            pix += bp;
            tiles += bt;
            i++;
        }
        long long nn = 0;
        BS_TRY(g_arena.begin(!g_debug));
        int rc = cfg.aff_dtype == BS_DTYPE_U8 ? stage1_batch<uint8_t>(P, batch, A, frags_out, node_base, &nn, s, ext_labels, ext_nlabels)
                                               : stage1_batch<float>(P, batch, A, frags_out, node_base, &nn, s, ext_labels, ext_nlabels);
        if (ext_labels) {
            BS_ARG(i >= P.owned.size(), "bs_stage1_from_labels: the owned blocks do not fit one batch (lower the number of blocks per call)");
        }
        g_arena.end();
        if (rc != BS_OK) return rc;
        node_base += nn;
    }
    P.n_nodes = node_base;
    g_prof.finish(s);
    // dense numbering over the blocks known so far (single rank: complete)
    P.block_nbase[0] = 0;
    for (size_t b = 0; b < P.blocks.size(); b++) P.block_nbase[b + 1] = P.block_nbase[b] + P.block_count[b];
    P.node_first = P.owned.empty() ? 0 : P.block_nbase[P.owned[0]];
    return BS_OK;
}

}  // namespace bs
