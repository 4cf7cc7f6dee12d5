// Device-wide primitives: exclusive scan and a stable LSD radix sort (hand-written, no CUB).
#include "common.cuh"

namespace bs {

static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
const char *get_error() { return g_err.c_str(); }
thread_local unsigned long long g_launches = 0;
thread_local Profiler g_prof;
thread_local Arena g_arena;

int Arena::begin(bool enable) {
    active = false;
    off = 0;
    need = 0;
    if (!enable) return BS_OK;
    int cur = 0;
    BS_CUDA(cudaGetDevice(&cur));
    if (base && cur != device) {
        // the thread moved to another device: the old slab stays where it is until it is released, a new one is sized here
        int old = device;
        cudaSetDevice(old);
        cudaDeviceSynchronize();
        cudaFree(base);
        cudaSetDevice(cur);
        base = nullptr;
        cap = 0;
    }
    device = cur;
    if (need_last > cap) {
        // nothing of the previous stage call is live any more: stage results are plan-owned (alloc_persistent)
        BS_CUDA(cudaDeviceSynchronize());
        if (base) BS_CUDA(cudaFree(base));
        base = nullptr;
        cap = 0;
        size_t want = need_last + need_last / 8 + (64u << 20);
        // the first run of this size went through the stream-ordered pool, which still caches those bytes
        {
            int dev = 0;
            cudaMemPool_t pool;
            if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
        }
        if (cudaMalloc((void **)&base, want) == cudaSuccess) {
            cap = want;
        } else {
            cudaGetLastError();   // not enough memory for a slab: stay on the stream-ordered allocator
            base = nullptr;
            need_last = 0;
        }
    }
    active = true;   // also without a slab: the bytes this run needs are recorded
    return BS_OK;
}

void Arena::end() {
    if (need > need_last) need_last = need;
    active = false;
}

void Arena::destroy() {
    if (base) cudaFree(base);
    base = nullptr;
    cap = off = need = need_last = 0;
    active = false;
}

void Profiler::mark(const char *name, cudaStream_t s) {
    if (!on) return;
    Ev e;
    e.name = name;
    cudaEventCreate(&e.ev);
    cudaEventRecord(e.ev, s);
    evs.push_back(e);
}
void Profiler::finish(cudaStream_t s) {
    if (!on || evs.empty()) return;
    mark("", s);
    cudaEventSynchronize(evs.back().ev);
    for (size_t i = 0; i + 1 < evs.size(); i++) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, evs[i].ev, evs[i + 1].ev);
        bool found = false;
        for (auto &r : result)
            if (r.first == evs[i].name) {
                r.second += ms;
                found = true;
            }
        if (!found) result.push_back(std::make_pair(evs[i].name, ms));
    }
    for (auto &e : evs) cudaEventDestroy(e.ev);
    evs.clear();
}
void Profiler::reset() {
    for (auto &e : evs) cudaEventDestroy(e.ev);
    evs.clear();
    result.clear();
}

// ------------------------------------------------------------------ scan
// 3-phase reduce-then-scan: tile = 256 threads x 16 items.
constexpr int SCAN_T = 256;
constexpr int SCAN_I = 16;
constexpr int SCAN_TILE = SCAN_T * SCAN_I;

template <typename TIn>
__global__ void scan_reduce_kernel(const TIn *__restrict__ in, uint32_t *__restrict__ tile_sums, size_t n) {
    __shared__ uint32_t wsum[SCAN_T / 32];
    size_t base = (size_t)blockIdx.x * SCAN_TILE;
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < SCAN_I; i++) {
        size_t idx = base + (size_t)i * SCAN_T + threadIdx.x;
        if (idx < n) acc += (uint32_t)in[idx];
    }
    acc = __reduce_add_sync(0xffffffffu, acc);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < SCAN_T / 32; w++) t += wsum[w];
        tile_sums[blockIdx.x] = t;
    }
}

// Each thread owns 16 CONSECUTIVE items so the in-thread order is the raster order.
template <typename TIn>
__global__ void scan_apply_kernel(const TIn *in, uint32_t *out, const uint32_t *__restrict__ tile_offsets,
                                  size_t n) {
    __shared__ uint32_t wsum[SCAN_T / 32];
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_I;
    uint32_t v[SCAN_I];
    uint32_t tsum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_I; i++) {
        size_t idx = base + i;
        v[i] = (idx < n) ? (uint32_t)in[idx] : 0u;
        tsum += v[i];
    }
    // warp inclusive scan of per-thread sums
    uint32_t inc = tsum;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < warp; w++) woff += wsum[w];
    uint32_t run = tile_offsets[blockIdx.x] + woff + (inc - tsum);
#pragma unroll
    for (int i = 0; i < SCAN_I; i++) {
        size_t idx = base + i;
        if (idx < n) out[idx] = run;
        run += v[i];
    }
}

__global__ void scan_small_kernel(uint32_t *data, size_t n, uint32_t *total) {
    // single CTA sequential-by-chunks scan for n <= a few thousand
    __shared__ uint32_t carry;
    __shared__ uint32_t wsum[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (size_t base = 0; base < n; base += blockDim.x) {
        size_t idx = base + threadIdx.x;
        uint32_t v = idx < n ? data[idx] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        uint32_t woff = 0;
        for (int w = 0; w < warp; w++) woff += wsum[w];
        uint32_t c = carry;
        if (idx < n) data[idx] = c + woff + inc - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = c + woff + inc;
        __syncthreads();
    }
    (void)nw;
    if (threadIdx.x == 0 && total) *total = carry;
}

template <typename TIn>
static int scan_impl(const TIn *in, uint32_t *out, size_t n, uint32_t *total_dev, cudaStream_t s) {
    if (n == 0) {
        if (total_dev) BS_CUDA(cudaMemsetAsync(total_dev, 0, 4, s));
        return BS_OK;
    }
    size_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    DevBuf sums;
    BS_TRY(sums.alloc(ntiles * 4, s));
    BS_LAUNCH((scan_reduce_kernel<TIn>), (unsigned)ntiles, SCAN_T, 0, s, in, sums.as<uint32_t>(), n);
    if (ntiles <= 8192) {
        BS_LAUNCH(scan_small_kernel, 1, 1024, 0, s, sums.as<uint32_t>(), ntiles, total_dev);
    } else {
        BS_TRY(scan_impl<uint32_t>(sums.as<uint32_t>(), sums.as<uint32_t>(), ntiles, total_dev, s));
    }
    BS_LAUNCH((scan_apply_kernel<TIn>), (unsigned)ntiles, SCAN_T, 0, s, in, out, sums.as<uint32_t>(), n);
    BS_CUDA(cudaGetLastError());
    return BS_OK;
}

int scan_exclusive_u32(const uint32_t *in, uint32_t *out, size_t n, uint32_t *total_dev, cudaStream_t s) {
    return scan_impl<uint32_t>(in, out, n, total_dev, s);
}
int scan_exclusive_u8(const uint8_t *in, uint32_t *out, size_t n, uint32_t *total_dev, cudaStream_t s) {
    return scan_impl<uint8_t>(in, out, n, total_dev, s);
}

// ------------------------------------------------------------------ radix sort
// One warp owns one tile of RS_TILE consecutive items and walks it 32 items at a time, so
// ranking inside a tile is trivially stable (__match_any_sync gives the same-digit lanes).
constexpr int RS_WARPS = 4;
constexpr int RS_STEPS = 64;
constexpr int RS_TILE = 32 * RS_STEPS;

__global__ void rs_hist_kernel(const uint64_t *__restrict__ keys, size_t n, int bit, uint32_t *__restrict__ hist,
                               size_t ntiles) {
    __shared__ uint32_t cnt[RS_WARPS][256];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    size_t tile = (size_t)blockIdx.x * RS_WARPS + warp;
    for (int d = lane; d < 256; d += 32) cnt[warp][d] = 0;
    __syncwarp();
    if (tile < ntiles) {
        size_t base = tile * RS_TILE;
        for (int st = 0; st < RS_STEPS; st++) {
            size_t i = base + (size_t)st * 32 + lane;
            if (i < n) {
                unsigned d = (unsigned)((keys[i] >> bit) & 255u);
                atomicAdd(&cnt[warp][d], 1u);
            }
        }
        __syncwarp();
        for (int d = lane; d < 256; d += 32) hist[(size_t)d * ntiles + tile] = cnt[warp][d];
    }
}

__global__ void rs_scatter_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals,
                                  uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, size_t n, int bit,
                                  const uint32_t *__restrict__ offs, size_t ntiles) {
    __shared__ uint32_t cnt[RS_WARPS][256];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    size_t tile = (size_t)blockIdx.x * RS_WARPS + warp;
    if (tile >= ntiles) return;
    for (int d = lane; d < 256; d += 32) cnt[warp][d] = offs[(size_t)d * ntiles + tile];
    __syncwarp();
    size_t base = tile * RS_TILE;
    for (int st = 0; st < RS_STEPS; st++) {
        size_t i = base + (size_t)st * 32 + lane;
        bool valid = i < n;
        unsigned act = __ballot_sync(0xffffffffu, valid);
        if (!act) break;
        if (valid) {
            uint64_t k = keys[i];
            uint32_t v = vals[i];
            unsigned d = (unsigned)((k >> bit) & 255u);
            unsigned m = __match_any_sync(act, d);
            unsigned rank = __popc(m & lanemask_lt());
            uint32_t pos = cnt[warp][d] + rank;
            __syncwarp(act);
            if (rank == 0) cnt[warp][d] += __popc(m);
            keys_out[pos] = k;
            vals_out[pos] = v;
        }
        __syncwarp();
    }
}

int radix_sort_pairs(uint64_t *keys, uint32_t *vals, uint64_t *keys_tmp, uint32_t *vals_tmp, size_t n, int bit_lo,
                     int bit_hi, cudaStream_t s) {
    if (n <= 1) return BS_OK;
    size_t ntiles = (n + RS_TILE - 1) / RS_TILE;
    DevBuf hist;
    BS_TRY(hist.alloc(256 * ntiles * 4, s));
    unsigned grid = cdiv(ntiles, RS_WARPS);
    uint64_t *kin = keys, *kout = keys_tmp;
    uint32_t *vin = vals, *vout = vals_tmp;
    int passes = 0;
    for (int bit = bit_lo; bit < bit_hi; bit += 8) {
        BS_LAUNCH(rs_hist_kernel, grid, RS_WARPS * 32, 0, s, kin, n, bit, hist.as<uint32_t>(), ntiles);
        BS_TRY(scan_exclusive_u32(hist.as<uint32_t>(), hist.as<uint32_t>(), 256 * ntiles, nullptr, s));
        BS_LAUNCH(rs_scatter_kernel, grid, RS_WARPS * 32, 0, s, kin, vin, kout, vout, n, bit, hist.as<uint32_t>(),
                  ntiles);
        uint64_t *tk = kin;
        kin = kout;
        kout = tk;
        uint32_t *tv = vin;
        vin = vout;
        vout = tv;
        passes++;
    }
    if (passes & 1) {
        BS_CUDA(cudaMemcpyAsync(keys, keys_tmp, n * 8, cudaMemcpyDeviceToDevice, s));
        BS_CUDA(cudaMemcpyAsync(vals, vals_tmp, n * 4, cudaMemcpyDeviceToDevice, s));
    }
    BS_CUDA(cudaGetLastError());
    return BS_OK;
}

}  // namespace bs
