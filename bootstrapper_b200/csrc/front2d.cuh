// Fused front end of stage 1 for 2-D tiles (the default fragments_in_xy mode, unshifted affinities).  Included by stage1.cu.
//
//   k_mask_bits*     affinities -> one bit per tile pixel (post/ws.py:64,77 boundary mask), row-padded bitmap.  u8 input:
//                    16-byte vector loads (LDG.128) of the a_y / a_x rows, or -- when the rows are 16-byte aligned in
//                    global memory -- TMA boxes (cp.async.bulk.tensor, zero fill outside the volume for free) staged
//                    through a shared-memory ring of bricks
//   k_tile_front     one CTA per tile, the whole tile on chip as 16-bit values: row distance -> exact squared EDT
//                    (scipy's distance_transform_edt incl. its all-foreground rule) -> seeds = plateau maxima of the
//                    min_seed_distance window (maximum_filter mode='reflect' == d2, post/ws.py:16-17) -> conn-1 seed
//                    components (scipy.ndimage.label, :19) -> dense priority levels + per-level FIFO segments.
//                    Leaves a 16-bit level per pixel, the floodable bitmap and the seed entries in HBM; the mask /
//                    row-distance / d2 / seed-parent planes of the unfused chain never exist
//                    (the flood itself is k_flood2 with FR = true: seeds arrive as ready-made queue entries, level tables in
//                    fixed slots per tile)
//
// Everything here computes the same integers as the unfused kernels (k_mask_rowdist, k_coldist_strip, k_maxfilt_xy_t,
// k_seed_union, k_seed_label_hist, k_levels, k_pixel_levels); tests/test_gpu_parity.py::test_front_versions_agree
// compares the two chains bit for bit.
// (no include guard / namespace: stage1.cu includes this file once, inside namespace bs, after the flood kernels)

static constexpr uint32_t FR_D2CAP = 16384;        // fused path: squared distances below this (radius < 128)
static constexpr int FR_NT = 1024;                 // threads per tile CTA
static constexpr int FR_MAXG = 4;                  // 4-row groups per warp and strip in the column pass: H <= 512
static constexpr int FR_SMEM_TOTAL = 232448 - 2048;   // dynamic shared memory of k_tile_front (static arrays + reserve aside)
static constexpr uint32_t FR_OVF_D2 = 1, FR_OVF_SEEDS = 2, FR_OVF_LEVELS = 4;

// ------------------------------------------------------------------ mask bitmap (row-padded: word (y, w) of tile t at
// mbase[t] + y * WW + w, WW = ceil(W / 32); bit j of a word = pixel x = 32 w + j)
// 16 bytes from an arbitrary address through aligned 16-byte loads (the two halves are neighbours in memory: the
// second load of one lane is the first load of the next, L1 serves it)
__device__ __forceinline__ uint4 ld16_unaligned(const uint8_t *p) {
    const uintptr_t a = (uintptr_t)p;
    const uint32_t r = (uint32_t)(a & 15);
    const uint4 *q = (const uint4 *)(a - r);
    const uint4 lo = __ldg(q);
    if (r == 0) return lo;
    const uint4 hi = __ldg(q + 1);
    const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    const uint32_t rw = r >> 2, sh = (r & 3) * 8;
    uint4 o;
    uint32_t v[5];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        // w[rw + k], rw in 0..3, selected without dynamic indexing into registers
        uint32_t x = w[k];
        if (rw == 1) x = w[k + 1];
        if (rw == 2) x = w[k + 2];
        if (rw == 3) x = w[k + 3 < 8 ? k + 3 : 7];
        v[k] = x;
    }
    o.x = __funnelshift_r(v[0], v[1], sh);
    o.y = __funnelshift_r(v[1], v[2], sh);
    o.z = __funnelshift_r(v[2], v[3], sh);
    o.w = __funnelshift_r(v[3], v[4], sh);
    return o;
}

// per-byte "a + b > 255" of two packed byte quads -> 4 bits
__device__ __forceinline__ uint32_t gt255_bits4(uint32_t a, uint32_t b) {
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) m |= ((((a >> (8 * k)) & 255u) + ((b >> (8 * k)) & 255u)) > 255u ? 1u : 0u) << k;
    return m;
}

// u8, 2-D tiles: warp per tile row, lane per 16 pixels.  vmask: optional (Z,Y,X) u8 volume mask.
__global__ void __launch_bounds__(256) k_mask_bits_u8(const Tile *__restrict__ tiles, const uint32_t *__restrict__ mbase, AffView A,
                                                      uint32_t *__restrict__ mbits) {
    const Tile t = tiles[blockIdx.y];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = t.W, H = t.H, WW = (W + 31) >> 5;
    const size_t nvol = (size_t)A.Zw * A.Y * A.X;
    const uint8_t *a = (const uint8_t *)A.p;
    uint32_t *out = mbits + mbase[blockIdx.y];
    const int gz = t.gz;
    const bool zin = gz >= 0 && gz < A.Z && gz >= A.z0 && gz < A.z0 + A.Zw;
    // four rows per warp and pass: their loads are independent and go out together
    for (int yq = (blockIdx.x * 8 + warp) * 4; yq < H; yq += gridDim.x * 32) {
#pragma unroll
      for (int yr = 0; yr < 4; yr++) {
        const int y = yq + yr;
        if (y >= H) break;
        const int gy = t.gy + y;
        const bool rowin = zin && gy >= 0 && gy < A.Y;
        const size_t rowoff = rowin ? ((size_t)(gz - A.z0) * A.Y + gy) * A.X : 0;
        for (int x0 = 0; x0 < W; x0 += 512) {
            const int x = x0 + lane * 16;          // first pixel of this lane
            uint32_t bits = 0;
            if (rowin && x < W) {
                const int gx = t.gx + x;
                // pixels [gx, gx + 16) clipped to the volume: valid bit range
                const int lo = max(0, -gx), hi = min(16, min(W - x, A.X - gx));
                if (hi > lo) {
                    uint4 vy, vx;
                    if (gx >= 0 && gx + 16 <= A.X) {
                        vy = ld16_unaligned(a + nvol + rowoff + gx);
                        vx = ld16_unaligned(a + 2 * nvol + rowoff + gx);
                    } else {
                        // row end / start inside the 16 pixels: byte loads
                        uint32_t wy[4] = {0, 0, 0, 0}, wx[4] = {0, 0, 0, 0};
                        for (int k = lo; k < hi; k++) {
                            wy[k >> 2] |= (uint32_t)a[nvol + rowoff + gx + k] << (8 * (k & 3));
                            wx[k >> 2] |= (uint32_t)a[2 * nvol + rowoff + gx + k] << (8 * (k & 3));
                        }
                        vy = make_uint4(wy[0], wy[1], wy[2], wy[3]);
                        vx = make_uint4(wx[0], wx[1], wx[2], wx[3]);
                    }
                    bits = gt255_bits4(vy.x, vx.x) | (gt255_bits4(vy.y, vx.y) << 4) | (gt255_bits4(vy.z, vx.z) << 8) |
                           (gt255_bits4(vy.w, vx.w) << 12);
                    bits &= ((hi >= 16 ? 0xFFFFu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u));
                    if (A.mask && bits) {
                        for (int k = lo; k < hi; k++)
                            if (((bits >> k) & 1u) && A.mask[rowoff + gx + k] == 0) bits &= ~(1u << k);
                    }
                }
            }
            // lanes (2j, 2j + 1) -> word j
            const uint32_t up = __shfl_down_sync(FULL, bits, 1);
            const int w = (x0 >> 5) + (lane >> 1);
            if ((lane & 1) == 0 && w < WW) out[(size_t)y * WW + w] = bits | (up << 16);
        }
      }
    }
}

// f32 (and the generic fallback): lane per pixel, ballot per word
template <typename T>
__global__ void __launch_bounds__(256) k_mask_bits_generic(const Tile *__restrict__ tiles, const uint32_t *__restrict__ mbase, AffView A,
                                                           uint32_t *__restrict__ mbits) {
    const Tile t = tiles[blockIdx.y];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = t.W, H = t.H, WW = (W + 31) >> 5;
    const size_t nvol = (size_t)A.Zw * A.Y * A.X;
    const T *a = (const T *)A.p;
    uint32_t *out = mbits + mbase[blockIdx.y];
    const int gz = t.gz;
    const bool zin = gz >= 0 && gz < A.Z && gz >= A.z0 && gz < A.z0 + A.Zw;
    for (int y = blockIdx.x * 8 + warp; y < H; y += gridDim.x * 8) {
        const int gy = t.gy + y;
        const bool rowin = zin && gy >= 0 && gy < A.Y;
        const size_t rowoff = rowin ? ((size_t)(gz - A.z0) * A.Y + gy) * A.X : 0;
        for (int c = 0; c < WW; c++) {
            const int x = c * 32 + lane, gx = t.gx + x;
            bool m = false;
            if (rowin && x < W && gx >= 0 && gx < A.X) {
                const size_t i = rowoff + gx;
                if (!A.mask || A.mask[i] > 0) m = AffOps<T>::boundary(a, nvol, i, 2);
            }
            const unsigned b = __ballot_sync(FULL, m);
            if (lane == 0) out[(size_t)y * WW + c] = b;
        }
    }
}

// ---- TMA variant (u8, rows 16-byte aligned in global memory: X % 16 == 0, base 16-byte aligned, no volume mask) -------
// One CTA walks a list of bricks = (tile, strip of TB_ROWS rows); a brick is two TMA boxes (channels a_y and a_x of the
// tile's z plane, TB_ROWS rows x `pitch` bytes) landing in a 2-stage shared-memory ring; voxels outside the volume arrive
// as zeros (the tile's read ROI reaches beyond the volume at its borders), which is exactly the reference's zero fill
// (to_ndarray(roi, fill_value=0), watershed_frags.py:197).
// Measured on B200 (tools/tma_probe.cu): the innermost box coordinate must be 16-byte aligned -- a misaligned start raises
// "illegal instruction", negative (aligned) starts are fine.  The tensor is therefore described as uint32 elements
// (X / 4 per row, which also lifts the 256-element box limit to 1024 bytes), boxes start at the 16-byte boundary at or below
// the tile's first column, and the consumer reads its pixels `shift` bytes into the row.
static constexpr int TB_ROWS = 16, TB_STAGES = 2;

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(phase)
            : "memory");
    }
}
// 4-D box load (x, y, z, c) of the (C, Z, Y, X / 4) uint32 view of the affinity tensor
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int x, int y, int z, int c) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            (uint32_t)__cvta_generic_to_shared(dst)),
        "l"((uint64_t)map), "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(x), "r"(y), "r"(z), "r"(c)
        : "memory");
}
// 16 bytes at byte offset o of a shared-memory row (rows are 16-byte aligned, the row holds o + 31 bytes at least)
__device__ __forceinline__ uint4 lds16_unaligned(const uint8_t *row, int o) {
    const uint32_t r = (uint32_t)o & 15u;
    const uint4 *q = (const uint4 *)(row + (o - (int)r));
    const uint4 lo = q[0];
    if (r == 0) return lo;
    const uint4 hi = q[1];
    const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    const uint32_t rw = r >> 2, sh = (r & 3) * 8;
    uint32_t v[5];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        uint32_t x = w[k];
        if (rw == 1) x = w[k + 1];
        if (rw == 2) x = w[k + 2];
        if (rw == 3) x = w[k + 3 < 8 ? k + 3 : 7];
        v[k] = x;
    }
    uint4 o4;
    o4.x = __funnelshift_r(v[0], v[1], sh);
    o4.y = __funnelshift_r(v[1], v[2], sh);
    o4.z = __funnelshift_r(v[2], v[3], sh);
    o4.w = __funnelshift_r(v[3], v[4], sh);
    return o4;
}

__global__ void __launch_bounds__(256) k_mask_bits_tma(const CUtensorMap *__restrict__ amap, const Tile *__restrict__ tiles,
                                                       const uint32_t *__restrict__ mbase, int ntiles, int strips_per_tile, int z0,
                                                       int pitch, uint32_t *__restrict__ mbits) {
    // brick = [2 channels][TB_ROWS][pitch] bytes per stage
    extern __shared__ __align__(128) uint8_t tb_smem[];
    __shared__ __align__(8) uint64_t bars[TB_STAGES];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nbricks = ntiles * strips_per_tile;
    if (threadIdx.x == 0) {
        for (int sidx = 0; sidx < TB_STAGES; sidx++) mbar_init(&bars[sidx], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t chan_bytes = (size_t)TB_ROWS * pitch, stage_bytes = 2 * chan_bytes;
    auto issue = [&](int brick, int stage) {
        const int ti = brick / strips_per_tile, st = brick - ti * strips_per_tile;
        const Tile t = tiles[ti];
        uint8_t *dst = tb_smem + (size_t)stage * stage_bytes;
        const int xe = (t.gx >> 4) * 4;       // arithmetic shift: floor for negative starts; in uint32 elements
        mbar_expect_tx(&bars[stage], (uint32_t)stage_bytes);
        tma_load_4d(dst, amap, &bars[stage], xe, t.gy + st * TB_ROWS, t.gz - z0, 1);
        tma_load_4d(dst + chan_bytes, amap, &bars[stage], xe, t.gy + st * TB_ROWS, t.gz - z0, 2);
    };
    int it = 0;
    if (threadIdx.x == 0 && (int)blockIdx.x < nbricks) issue(blockIdx.x, 0);
    for (int brick = blockIdx.x; brick < nbricks; brick += gridDim.x, it++) {
        const int stage = it & 1;
        const int next = brick + gridDim.x;
        if (threadIdx.x == 0 && next < nbricks) issue(next, stage ^ 1);
        mbar_wait(&bars[stage], (uint32_t)((it >> 1) & 1));
        const int ti = brick / strips_per_tile, st = brick - ti * strips_per_tile;
        const Tile t = tiles[ti];
        const int W = t.W, WW = (W + 31) >> 5;
        const int shift = t.gx - ((t.gx >> 4) << 4);
        const uint8_t *sy = tb_smem + (size_t)stage * stage_bytes;   // [row][pitch] of channel a_y
        const uint8_t *sx = sy + chan_bytes;                         // channel a_x
        uint32_t *out = mbits + mbase[ti];
        // 8 warps x 2 rows; lane per 16 pixels
        for (int r = warp; r < TB_ROWS; r += 8) {
            const int y = st * TB_ROWS + r;
            for (int x0 = 0; x0 < W; x0 += 512) {
                const int x = x0 + lane * 16;
                uint32_t bits = 0;
                if (y < t.H && x < W) {
                    const uint4 vy = lds16_unaligned(sy + (size_t)r * pitch, x + shift);
                    const uint4 vx = lds16_unaligned(sx + (size_t)r * pitch, x + shift);
                    bits = gt255_bits4(vy.x, vx.x) | (gt255_bits4(vy.y, vx.y) << 4) | (gt255_bits4(vy.z, vx.z) << 8) |
                           (gt255_bits4(vy.w, vx.w) << 12);
                    if (W - x < 16) bits &= (1u << (W - x)) - 1u;
                }
                const uint32_t up = __shfl_down_sync(FULL, bits, 1);
                const int w = (x0 >> 5) + (lane >> 1);
                if ((lane & 1) == 0 && w < WW && y < t.H) out[(size_t)y * WW + w] = bits | (up << 16);
            }
        }
        __syncthreads();   // the stage is free again before its next brick is issued (two iterations ahead)
    }
}

// ------------------------------------------------------------------ the tile front end
// row pitch (in 16-bit elements) of the on-chip planes: an odd number of 32-bit words, so that lanes walking down a
// column of rows hit different banks
__host__ __device__ __forceinline__ int fr_pitch(int W) {
    int p = (W + 1) & ~1;
    if (((p >> 1) & 1) == 0) p += 2;
    return p;
}

__device__ __forceinline__ uint32_t block_excl_scan_1024(uint32_t v, uint32_t *wsum /*[33] shared*/, uint32_t &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t2 = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t2;
    }
    __syncthreads();
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = wsum[lane], winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t2 = __shfl_up_sync(FULL, winc, o);
            if (lane >= o) winc += t2;
        }
        wsum[lane] = winc - w;
        if (lane == 31) wsum[32] = winc;
    }
    __syncthreads();
    total = wsum[32];
    return wsum[warp] + inc - v;
}

__device__ __forceinline__ uint32_t uf_find_s(volatile uint32_t *par, uint32_t x) {
    uint32_t p = par[x];
    while (p != x) {
        x = p;
        p = par[x];
    }
    return x;
}
__device__ __forceinline__ void uf_union_s(uint32_t *par, uint32_t a, uint32_t b) {
    for (;;) {
        a = uf_find_s(par, a);
        b = uf_find_s(par, b);
        if (a == b) return;
        if (a < b) {
            const uint32_t t2 = a;
            a = b;
            b = t2;
        }
        const uint32_t old = atomicMin(&par[a], b);
        if (old == a) return;
        a = old;
    }
}

struct FrontOut {
    uint16_t *lv16;         // per tile pixel: dense priority level (rank of d2 among the tile's values; 0 outside the mask)
    uint32_t *availw;       // per tile pixel, one bit: floodable (in mask, not a seed); word (t.base >> 5) + i / 32
    uint32_t *seedent;      // [t.base + k]: (label << 17 | pixel) of the tile's k-th seed pixel in raster order
    uint32_t *lvl_qstart;   // [tile * levtab + r]: queue position where the FIFO of the tile's level r starts
    uint32_t *lvl_head;     // same indexing, zeroed here
    uint32_t *t_nseeds, *t_nlev, *t_nmask;
    uint32_t *flags;        // [0] overflow bits, [1] max levels per tile, [2] max seed pixels per tile, [3] any d2 > 0
    int levtab;
};

__global__ void __launch_bounds__(FR_NT, 1) k_tile_front(const Tile *__restrict__ tiles, const uint32_t *__restrict__ mbase,
                                                         const uint32_t *__restrict__ mbits, int msd, int scr_bytes, FrontOut O) {
    extern __shared__ __align__(16) uint8_t fr_smem[];
    __shared__ uint32_t wsum[33];
    __shared__ uint32_t s_flags, s_tmax;
    const Tile t = tiles[blockIdx.x];
    const int W = t.W, H = t.H, WW = (W + 31) >> 5, Wp = fr_pitch(W);
    const int npix = H * W, nwords = H * WW;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint16_t *D = (uint16_t *)fr_smem;                                    // [H][Wp] row distance, then squared EDT
    uint32_t *B = (uint32_t *)(fr_smem + (((size_t)H * Wp * 2 + 15) & ~(size_t)15));   // [H][WW] mask bits, then seed bits
    uint8_t *SCR = (uint8_t *)(B + nwords);
    // the last 3 KB of the scratch area live through all phases: bitmap of the d2 values present in the mask (filled by the
    // column pass) and its prefix popcounts; the rest is phase-local (maximum-filter band, seed ranks / union-find, level counts)
    const int scr_work = scr_bytes - 3072;
    uint32_t *pres = (uint32_t *)(SCR + scr_work);              // [512] bitmap of the d2 values present
    uint16_t *ppre = (uint16_t *)(SCR + scr_work + 2048);       // [512] values present before the word
    for (int i = tid; i < 512; i += FR_NT) pres[i] = 0;
    if (tid == 0) s_flags = 0, s_tmax = 0;
    // ---- P0: mask bits; the pad columns of D stay 0 (= outside the mask) throughout
    for (int i = tid; i < nwords; i += FR_NT) B[i] = mbits[(size_t)mbase[blockIdx.x] + i];
    for (int i = tid; i < H * (Wp - W); i += FR_NT) D[(i / (Wp - W)) * Wp + W + i % (Wp - W)] = 0;
    __syncthreads();
    // ---- P1: distance to the nearest background pixel along x (GINF: none in the row).  Lane c first holds word c of the
    //          row (W <= 1024): the last background pixel before each word / the first after it follow from a prefix
    //          maximum / suffix minimum over the lanes, then every pixel needs its own word and those two positions
    for (int y = warp; y < H; y += FR_NT / 32) {
        const uint32_t *row = B + y * WW;
        uint32_t myw = 0;
        if (lane < WW) {
            const int rem = W - lane * 32;
            myw = ~row[lane] & (rem >= 32 ? FULL : ((1u << rem) - 1u));     // background bits
        }
        // position + 1 of the last background pixel in words <= lane (0: none); first in words >= lane (0x7FFFFFFF: none)
        int last = myw ? lane * 32 + 32 - __clz(myw) : 0;
        int first = myw ? lane * 32 + __ffs(myw) - 1 : 0x7FFFFFFF;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int a = __shfl_up_sync(FULL, last, o), bq = __shfl_down_sync(FULL, first, o);
            if (lane >= o) last = max(last, a);
            if (lane + o < 32) first = min(first, bq);
        }
        int last_prev = __shfl_up_sync(FULL, last, 1), first_next = __shfl_down_sync(FULL, first, 1);
        if (lane == 0) last_prev = 0;
        if (lane == 31) first_next = 0x7FFFFFFF;
        for (int c = 0; c < WW; c++) {
            const int x = c * 32 + lane;
            const uint32_t word = __shfl_sync(FULL, myw, c);
            const int lp = __shfl_sync(FULL, last_prev, c), fn = __shfl_sync(FULL, first_next, c);
            if (x < W) {
                uint32_t dist = 0;
                if (!((word >> lane) & 1u)) {
                    const uint32_t wl = word & ((1u << lane) - 1u);
                    const uint32_t wr = word & ~((2u << lane) - 1u);
                    const int pl = wl ? c * 32 + 32 - __clz(wl) : lp;                 // position + 1, 0: none
                    const int pr = wr ? c * 32 + __ffs(wr) - 1 : fn;                  // position, 0x7FFFFFFF: none
                    const uint32_t dl = pl ? (uint32_t)(x - (pl - 1)) : (uint32_t)GINF;
                    const uint32_t dr = pr != 0x7FFFFFFF ? (uint32_t)(pr - x) : (uint32_t)GINF;
                    dist = min(dl, dr);
                }
                D[y * Wp + x] = (uint16_t)dist;
            }
        }
    }
    __syncthreads();
    // ---- P2: exact squared EDT, column pass in strips of 32 columns (results staged in registers, written in place)
    {
        uint32_t mymax = 0, myflags = 0;
        const int ngroups = (H + 3) >> 2;
        for (int x0 = 0; x0 < W; x0 += 32) {
            const int x = x0 + lane;
            uint32_t res[FR_MAXG][4];
            if (x < W) {
                const uint16_t *col = D + x;
#pragma unroll
                for (int gi = 0; gi < FR_MAXG; gi++) {
                    const int grp = warp + gi * (FR_NT / 32);
                    if (grp >= ngroups) break;
                    const int ya = grp * 4;
                    uint32_t g0[4], b[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        uint32_t v = ya + j < H ? col[(ya + j) * Wp] : GINF;
                        g0[j] = v == GINF ? DBIG : v * v;
                    }
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        b[j] = g0[j];
#pragma unroll
                        for (int i2 = 0; i2 < 4; i2++)
                            if (i2 != j) b[j] = min(b[j], g0[i2] + (uint32_t)((i2 - j) * (i2 - j)));
                    }
                    // rows of the group beyond the tile never win: their pixels are not stored
#pragma unroll
                    for (int j = 1; j < 4; j++)
                        if (ya + j >= H) b[j] = 0;
                    for (int k = 1;; k++) {
                        const uint32_t kk = (uint32_t)k * k;
                        const int ru = ya - k, rd = ya + 3 + k;
                        const uint32_t bm = max(max(b[0], b[1]), max(b[2], b[3]));
                        if (kk >= bm || (ru < 0 && rd >= H)) break;
                        const uint32_t k2 = 2u * (uint32_t)k;
                        if (ru >= 0) {
                            const uint32_t v = col[ru * Wp];
                            const uint32_t g2 = (v == GINF ? DBIG : v * v) + kk;
                            b[0] = min(b[0], g2);
                            b[1] = min(b[1], g2 + k2 + 1u);
                            b[2] = min(b[2], g2 + 2u * k2 + 4u);
                            b[3] = min(b[3], g2 + 3u * k2 + 9u);
                        }
                        if (rd < H) {
                            const uint32_t v = col[rd * Wp];
                            const uint32_t g2 = (v == GINF ? DBIG : v * v) + kk;
                            b[3] = min(b[3], g2);
                            b[2] = min(b[2], g2 + k2 + 1u);
                            b[1] = min(b[1], g2 + 2u * k2 + 4u);
                            b[0] = min(b[0], g2 + 3u * k2 + 9u);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        uint32_t v = b[j];
                        // scipy's all-foreground rule for a 2-D array: the background sits at (-1, 0)
                        if (v >= DBIG) v = (uint32_t)(ya + j + 1) * (ya + j + 1) + (uint32_t)x * x;
                        res[gi][j] = v;
                    }
                }
            }
            __syncthreads();
            if (x < W) {
#pragma unroll
                for (int gi = 0; gi < FR_MAXG; gi++) {
                    const int grp = warp + gi * (FR_NT / 32);
                    if (grp >= ngroups) break;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int y = grp * 4 + j;
                        if (y < H) {
                            const uint32_t v = res[gi][j];
                            if (v >= FR_D2CAP)
                                myflags |= FR_OVF_D2;
                            else if (v && !((pres[v >> 5] >> (v & 31)) & 1u))
                                atomicOr(&pres[v >> 5], 1u << (v & 31));
                            mymax = max(mymax, v);
                            D[y * Wp + x] = (uint16_t)min(v, 65535u);
                        }
                    }
                }
            }
            // the next strip reads other columns only: no barrier needed before it starts
        }
        mymax = __reduce_max_sync(FULL, mymax);
        myflags = __reduce_or_sync(FULL, myflags);
        if (lane == 0) {
            if (mymax) atomicMax(&s_tmax, mymax);
            if (myflags) atomicOr(&s_flags, myflags);
        }
    }
    __syncthreads();
    // ---- P3: seeds = (maximum_filter(d2, msd, mode='reflect') == d2) & mask, separable: bands of rows whose x-filtered rows
    //          (msd - 1 extra rows, reflected at the tile border) sit in the scratch area, then the y pass per pixel.
    //          msd == 10 (the default): a thread produces G = 8 (or 4, when the scratch area holds fewer rows) neighbouring
    //          maxima from G + 9 inputs by doubling (max over 2, 4, 8, then 8 + 2), lanes walk down rows (x pass) / along
    //          rows (y pass): no bank conflicts
    {
        const int lo = msd / 2;
        uint16_t *R = (uint16_t *)SCR;
        const int rows_fit = scr_work / (2 * Wp);
        const int RB = ((rows_fit - (msd - 1)) / 4) * 4;           // rows per band (the host guarantees RB >= 4)
        for (int yb = 0; yb < H; yb += RB) {
            const int nb = min(RB, H - yb), nrows = nb + msd - 1;
            if (msd == 10) {
                const int nseg = (W + 7) >> 3, nrows32 = (nrows + 31) & ~31;
                for (int item = tid; item < nseg * nrows32; item += FR_NT) {
                    const int sg = item / nrows32, j = item - sg * nrows32;
                    if (j >= nrows) continue;
                    const int x0 = sg * 8;
                    const uint16_t *row = D + reflect_idx(yb - lo + j, H) * Wp;
                    uint32_t in[17];
                    if (x0 - lo >= 0 && x0 - lo + 17 <= W) {
#pragma unroll
                        for (int k = 0; k < 17; k++) in[k] = row[x0 - lo + k];
                    } else {
#pragma unroll
                        for (int k = 0; k < 17; k++) in[k] = row[reflect_idx(x0 - lo + k, W)];
                    }
                    uint32_t m2[16], m4[14], m8[10];
#pragma unroll
                    for (int k = 0; k < 16; k++) m2[k] = max(in[k], in[k + 1]);
#pragma unroll
                    for (int k = 0; k < 14; k++) m4[k] = max(m2[k], m2[k + 2]);
#pragma unroll
                    for (int k = 0; k < 10; k++) m8[k] = max(m4[k], m4[k + 4]);
                    uint16_t *dst = R + j * Wp + x0;
#pragma unroll
                    for (int k = 0; k < 8; k++)
                        if (x0 + k < W) dst[k] = (uint16_t)max(m8[k], m2[k + 8]);
                }
                __syncthreads();
                // groups of 4 output rows from 13 x-filtered rows
                const int ngrp = (nb + 3) >> 2;
                for (int wi = warp; wi < ngrp * WW; wi += FR_NT / 32) {
                    const int rg = wi / WW, c = wi - rg * WW;
                    const int x = c * 32 + lane;
                    uint32_t in[13];
                    const uint16_t *colp = R + (rg * 4) * Wp + min(x, W - 1);
#pragma unroll
                    for (int k = 0; k < 13; k++) in[k] = colp[k * Wp];     // rows past the band: stale values, results unused
                    uint32_t m2[12], m4[10], m8[6];
#pragma unroll
                    for (int k = 0; k < 12; k++) m2[k] = max(in[k], in[k + 1]);
#pragma unroll
                    for (int k = 0; k < 10; k++) m4[k] = max(m2[k], m2[k + 2]);
#pragma unroll
                    for (int k = 0; k < 6; k++) m8[k] = max(m4[k], m4[k + 4]);
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int yr = rg * 4 + k;
                        bool seed = false;
                        if (x < W && yr < nb) {
                            const uint32_t v = D[(yb + yr) * Wp + x];
                            seed = v != 0 && max(m8[k], m2[k + 8]) == v;
                        }
                        const unsigned sb = __ballot_sync(FULL, seed);
                        if (lane == 0 && yr < nb) B[(yb + yr) * WW + c] = sb;   // the mask bits are not needed any more (mask == d2 > 0)
                    }
                }
            } else {
                for (int i = tid; i < nrows * W; i += FR_NT) {
                    const int j = i / W, x = i - j * W;
                    const uint16_t *row = D + reflect_idx(yb - lo + j, H) * Wp;
                    uint32_t m = 0;
                    for (int k = 0; k < msd; k++) m = max(m, (uint32_t)row[reflect_idx(x - lo + k, W)]);
                    R[j * Wp + x] = (uint16_t)m;
                }
                __syncthreads();
                for (int wi = warp; wi < nb * WW; wi += FR_NT / 32) {
                    const int yr = wi / WW, c = wi - yr * WW;
                    const int x = c * 32 + lane;
                    bool seed = false;
                    if (x < W) {
                        const uint32_t v = D[(yb + yr) * Wp + x];
                        if (v) {
                            uint32_t m = 0;
                            for (int k = 0; k < msd; k++) m = max(m, (uint32_t)R[(yr + k) * Wp + x]);
                            seed = m == v;
                        }
                    }
                    const unsigned sb = __ballot_sync(FULL, seed);
                    if (lane == 0) B[(yb + yr) * WW + c] = sb;
                }
            }
            __syncthreads();
        }
    }
    __syncthreads();
    // ---- P4: raster ranks of the seed pixels: seeds before every row (a word's own offset is summed over the row's few words)
    uint16_t *rpre = (uint16_t *)SCR;                                             // [H] seeds before the row
    uint32_t *par = (uint32_t *)(SCR + (((size_t)H * 2 + 15) & ~(size_t)15));     // [seedcap] union-find over seed ranks
    const int seedcap = (int)((scr_work - (((size_t)H * 2 + 15) & ~(size_t)15)) / 4);
    uint32_t nseeds = 0;
    {
        uint32_t c = 0;
        if (tid < H)
            for (int w = 0; w < WW; w++) c += __popc(B[tid * WW + w]);
        const uint32_t run = block_excl_scan_1024(c, wsum, nseeds);     // H <= 512 rows, one per thread
        if (tid < H) rpre[tid] = (uint16_t)min(run, 65535u);
    }
    const bool seeds_ok = nseeds <= (uint32_t)seedcap && nseeds < 32767u;
    if (!seeds_ok && tid == 0) atomicOr(&s_flags, FR_OVF_SEEDS);
    __syncthreads();
    auto word_base = [&](int y, int c) -> uint32_t {      // seeds before word (y, c)
        uint32_t b = rpre[y];
        for (int w = 0; w < c; w++) b += __popc(B[y * WW + w]);
        return b;
    };
    // ---- P5: conn-1 components of the seed pixels; label = raster rank of the component's first pixel + 1
    if (seeds_ok) {
        for (uint32_t r = tid; r < nseeds; r += FR_NT) par[r] = r;
        __syncthreads();
        for (int wi = tid; wi < nwords; wi += FR_NT) {
            uint32_t word = B[wi];
            if (!word) continue;
            const int y = wi / WW, c = wi - y * WW;
            const uint32_t up = y > 0 ? B[wi - WW] : 0u;
            const uint32_t base = word_base(y, c);
            const uint32_t upbase = (up & word) ? word_base(y - 1, c) : 0u;   // only needed where a seed sits above a seed
            uint32_t rem = word;
            while (rem) {
                const int bpos = __ffs(rem) - 1;
                rem &= rem - 1;
                const uint32_t r = base + __popc(word & ((1u << bpos) - 1u));
                // left neighbour
                if (bpos > 0) {
                    if ((word >> (bpos - 1)) & 1u) uf_union_s(par, r, r - 1);
                } else if (c > 0 && (B[wi - 1] >> 31)) {
                    uf_union_s(par, r, r - 1);     // the previous seed in raster order is the pixel to the left
                }
                if ((up >> bpos) & 1u) uf_union_s(par, r, upbase + __popc(up & ((1u << bpos) - 1u)));
            }
        }
        __syncthreads();
        uint32_t *sent = O.seedent + t.base;
        for (int wi = tid; wi < nwords; wi += FR_NT) {
            const uint32_t word = B[wi];
            if (!word) continue;
            const int y = wi / WW, c = wi - y * WW;
            const uint32_t base = word_base(y, c);
            uint32_t rem = word;
            while (rem) {
                const int bpos = __ffs(rem) - 1;
                rem &= rem - 1;
                const uint32_t r = base + __popc(word & ((1u << bpos) - 1u));
                const uint32_t root = uf_find_s(par, r);
                sent[r] = ((root + 1u) << 17) | (uint32_t)(y * W + c * 32 + bpos);
            }
        }
    }
    __syncthreads();
    // ---- P6: dense priority levels (rank of d2 among the values present in the mask), per-level counts -> FIFO segments
    uint32_t *cnt = (uint32_t *)SCR;                        // [levcap] mask pixels per level
    const int levcap = scr_work / 4;
    for (int i = tid; i < levcap; i += FR_NT) cnt[i] = 0;
    const bool d2_ok = !(s_flags & FR_OVF_D2);
    __syncthreads();
    uint32_t nlev = 0;
    {
        const uint32_t c = tid < 512 ? __popc(pres[tid]) : 0u;
        const uint32_t ex = block_excl_scan_1024(c, wsum, nlev);
        if (tid < 512) ppre[tid] = (uint16_t)ex;
    }
    const bool lev_ok = d2_ok && nlev <= (uint32_t)levcap && nlev <= (uint32_t)O.levtab && nlev < 32768u;
    if (!lev_ok && d2_ok && tid == 0) atomicOr(&s_flags, FR_OVF_LEVELS);
    __syncthreads();
    uint16_t *lv16 = O.lv16 + t.base;
    uint32_t *availw = O.availw + (t.base >> 5);       // tile bases are 32-pixel aligned
    if (lev_ok && seeds_ok) {
        const int nchunks = (npix + 31) >> 5;
        for (int ch = warp; ch < nchunks; ch += FR_NT / 32) {
            const int i = ch * 32 + lane;
            bool av = false;
            if (i < npix) {
                const int y = (int)fdiv((uint32_t)i, t.fW), x = i - y * W;
                const uint32_t v = D[y * Wp + x];
                uint32_t r = 0;
                if (v) {
                    r = (uint32_t)ppre[v >> 5] + __popc(pres[v >> 5] & ((1u << (v & 31)) - 1u));
                    atomicAdd(&cnt[r], 1u);
                    av = !((B[y * WW + (x >> 5)] >> (x & 31)) & 1u);
                }
                lv16[i] = (uint16_t)r;
            }
            const unsigned aw = __ballot_sync(FULL, av);
            if (lane == 0) availw[ch] = aw;
        }
    }
    __syncthreads();
    uint32_t nmask = 0;
    if (lev_ok && seeds_ok) {
        const int per = ((int)nlev + FR_NT - 1) / FR_NT;
        const int l0 = tid * per;
        uint32_t c = 0;
        for (int i = 0; i < per; i++)
            if (l0 + i < (int)nlev) c += cnt[l0 + i];
        uint32_t run = block_excl_scan_1024(c, wsum, nmask);
        for (int i = 0; i < per; i++)
            if (l0 + i < (int)nlev) {
                O.lvl_qstart[(size_t)blockIdx.x * O.levtab + l0 + i] = (uint32_t)t.base + run;
                O.lvl_head[(size_t)blockIdx.x * O.levtab + l0 + i] = 0u;
                run += cnt[l0 + i];
            }
    }
    if (tid == 0) {
        O.t_nseeds[blockIdx.x] = seeds_ok ? nseeds : 0u;
        O.t_nlev[blockIdx.x] = lev_ok ? nlev : 0u;
        O.t_nmask[blockIdx.x] = nmask;
        if (s_flags) atomicOr(&O.flags[0], s_flags);
        atomicMax(&O.flags[1], nlev);
        atomicMax(&O.flags[2], nseeds);
    }
}

// labels leave the queue through shared memory (k_scatter_labels_tile with the tile's queue range given per tile)
__global__ void __launch_bounds__(SCAT_NT) k_scatter_labels_tile5(const Tile *__restrict__ tiles, const uint32_t *__restrict__ t_nmask,
                                                                  const uint32_t *__restrict__ queue, uint16_t *__restrict__ lab) {
    extern __shared__ uint16_t sc_lab[];
    const Tile t = tiles[blockIdx.x];
    const int npix = t.H * t.W;
    for (int i = threadIdx.x; i < (npix + 1) / 2; i += SCAT_NT) ((uint32_t *)sc_lab)[i] = 0u;
    __syncthreads();
    const uint32_t qb = (uint32_t)t.base, qe = qb + t_nmask[blockIdx.x];
    for (uint32_t q = qb + threadIdx.x; q < qe; q += SCAT_NT) {
        const uint32_t e = __ldcs(&queue[q]);
        if (e != NONE32) sc_lab[e & F2_PIXMASK] = (uint16_t)(e >> 17);
    }
    __syncthreads();
    uint32_t *out = (uint32_t *)(lab + t.base);        // tile bases are 32-pixel aligned: whole words
    for (int i = threadIdx.x; i < (npix + 1) / 2; i += SCAT_NT) out[i] = ((const uint32_t *)sc_lab)[i];
}

