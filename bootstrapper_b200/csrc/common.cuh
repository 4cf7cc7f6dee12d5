// Shared declarations for libbsnative (sm_100a).  Internal header, not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <utility>
#include <vector>

namespace bs {

// ---- error plumbing (thread-local message, C ABI returns negative codes) ----
void set_error(const std::string &msg);
const char *get_error();
#ifndef BS_OK
#define BS_OK 0
#define BS_ERR_CUDA (-1)
#define BS_ERR_ARG (-2)
#define BS_ERR_OVERFLOW (-3)
#define BS_ERR_STATE (-4)
#endif

#define BS_CUDA(call)                                                                         \
    do {                                                                                      \
        cudaError_t _e = (call);                                                              \
        if (_e != cudaSuccess) {                                                              \
            char _b[512];                                                                     \
            snprintf(_b, sizeof(_b), "%s:%d: %s -> %s", __FILE__, __LINE__, #call,            \
                     cudaGetErrorString(_e));                                                 \
            bs::set_error(_b);                                                                \
            return BS_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

#define BS_TRY(call)                 \
    do {                             \
        int _r = (call);             \
        if (_r != BS_OK) return _r;  \
    } while (0)

#define BS_ARG(cond, msg)            \
    do {                             \
        if (!(cond)) {               \
            bs::set_error(msg);      \
            return BS_ERR_ARG;       \
        }                            \
    } while (0)

// launch counter (bench.py reports gpu_launches from it)
extern thread_local unsigned long long g_launches;   // per host thread, like the arena and the profiler
#define BS_LAUNCH(kernel, grid, block, smem, stream, ...)          \
    do {                                                           \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); \
        bs::g_launches++;                                          \
    } while (0)

// ---- scratch memory ----
// Stage scratch comes from one grow-only arena per process (bump allocation, reset at the start of every stage
// batch): a stage issues ~50 allocations whose sizes depend on the data, and carving them from a cached slab keeps
// cudaMallocAsync / cudaFreeAsync (and the pool's occasional re-mapping) out of the step.  The first run of a
// given size finds no arena, allocates stream-ordered and records the bytes it needed; the next run gets the slab.
// Scratch slab of one host thread on one device: a stage call bump-allocates its temporaries from it and resets it when the
// next stage call of the same thread begins (stage calls synchronise their stream before they return).  thread_local: two
// host threads driving two plans never share scratch; a thread that switches devices gets a fresh slab.
struct Arena {
    char *base = nullptr;
    size_t cap = 0, off = 0, need = 0, need_last = 0;
    bool active = false;
    int device = -1;
    ~Arena() { if (base) cudaFree(base); }
    int begin(bool enable);   // (re)size from the recorded need, reset the bump pointer
    void end();
    void destroy();
};
extern thread_local Arena g_arena;

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    cudaStream_t s = 0;
    bool from_arena = false;
    bool persistent = false;   // plan-owned results: freed with cudaFree (the stream they were made on may be gone by then)
    DevBuf() {}
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    int alloc(size_t n, cudaStream_t stream) {
        release();
        s = stream;
        bytes = n;
        if (n == 0) n = 16;
        if (g_arena.active) {
            size_t a = (n + 255) & ~(size_t)255;
            g_arena.need += a;
            if (g_arena.off + a <= g_arena.cap) {
                p = g_arena.base + g_arena.off;
                g_arena.off += a;
                from_arena = true;
                return BS_OK;
            }
        }
        BS_CUDA(cudaMallocAsync(&p, n, stream));
        return BS_OK;
    }
    // results that outlive the stage call (plan-owned node / edge tables)
    int alloc_persistent(size_t n, cudaStream_t stream) {
        release();
        s = stream;
        bytes = n;
        if (n == 0) n = 16;
        BS_CUDA(cudaMallocAsync(&p, n, stream));
        persistent = true;
        return BS_OK;
    }
    int alloc_zero(size_t n, cudaStream_t stream) {
        BS_TRY(alloc(n, stream));
        BS_CUDA(cudaMemsetAsync(p, 0, n ? n : 16, stream));
        return BS_OK;
    }
    int alloc_fill(size_t n, int byte, cudaStream_t stream) {
        BS_TRY(alloc(n, stream));
        BS_CUDA(cudaMemsetAsync(p, byte, n ? n : 16, stream));
        return BS_OK;
    }
    void release() {
        if (p && !from_arena) {
            if (persistent)
                cudaFree(p);
            else
                cudaFreeAsync(p, s);
        }
        p = nullptr;
        bytes = 0;
        from_arena = false;
        persistent = false;
    }
    void swap(DevBuf &o) {
        std::swap(p, o.p);
        std::swap(bytes, o.bytes);
        std::swap(s, o.s);
        std::swap(from_arena, o.from_arena);
        std::swap(persistent, o.persistent);
    }
    template <typename T>
    T *as() const {
        return (T *)p;
    }
};

static inline unsigned int cdiv(size_t a, size_t b) { return (unsigned int)((a + b - 1) / b); }

// ---- per-phase device timing (CUDA events on the caller's stream), off by default ----
struct Profiler {
    bool on = false;
    struct Ev {
        std::string name;
        cudaEvent_t ev;
    };
    std::vector<Ev> evs;
    std::vector<std::pair<std::string, float>> result;
    void mark(const char *name, cudaStream_t s);   // marks the START of phase `name`
    void finish(cudaStream_t s);                   // closes the last phase, synchronises, accumulates
    void reset();
};
extern thread_local Profiler g_prof;

// ---- primitives (prims.cu) ----
// out[i] = sum_{j<i} in[j]; total (device pointer, may be null) = sum of all.  in/out may alias.
int scan_exclusive_u32(const uint32_t *in, uint32_t *out, size_t n, uint32_t *total_dev, cudaStream_t s);
int scan_exclusive_u8(const uint8_t *in, uint32_t *out, size_t n, uint32_t *total_dev, cudaStream_t s);
// stable LSD radix sort of (key u64, value u32) pairs on bits [bit_lo, bit_hi); result ends in keys/vals
// (tmp buffers of the same size are used for ping-pong).
int radix_sort_pairs(uint64_t *keys, uint32_t *vals, uint64_t *keys_tmp, uint32_t *vals_tmp, size_t n,
                     int bit_lo, int bit_hi, cudaStream_t s);

// mutex watershed (mws.cu)
int mws_agglom(const void *affs, int aff_dtype, const uint8_t *mask, int C, int Z, int Y, int X, const int32_t *offsets,
               const int32_t *strides, const double *bias, double noise_eps, unsigned long long noise_seed, int zero_is_repulsive,
               int remove_debris, uint64_t *labels_out, uint64_t *seg_out, int64_t *counters_out, cudaStream_t s);
int mws_agglom_blocks(const void *affs, int aff_dtype, const uint8_t *mask, int C, int n_blocks, int Z, int Y, int X, const int32_t *offsets,
                      const int32_t *strides, const double *bias, double noise_eps, const unsigned long long *block_seeds_host,
                      int zero_is_repulsive, uint32_t *labels_out, int64_t *counters_out, cudaStream_t s);
int graph_mws(const uint64_t *nodes, int64_t n, const uint64_t *u, const uint64_t *v, const float *scores, int64_t m, double weight,
              double bias, uint64_t *out, int64_t *counters_out, cudaStream_t s);

// ---- device helpers ----
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// division of 32-bit indices by a run-time constant (tile / block extents): multiply-high with a precomputed magic
// number (round-up method, exact for every 32-bit numerator)
struct FastDiv {
    uint32_t m, sh;   // m == 0: the divisor is a power of two, shift by sh
};
static inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    uint32_t l = 0;
    while ((1ull << l) < d) l++;
    if ((1ull << l) == d) {
        f.m = 0;
        f.sh = l;
    } else {
        f.m = (uint32_t)((((1ull << l) - d) << 32) / d) + 1;
        f.sh = l - 1;
    }
    return f;
}
// the same magic number computed on the device (one 64-bit division per call: hoist it out of loops)
__device__ __forceinline__ FastDiv make_fastdiv_dev(uint32_t d) {
    FastDiv f;
    const uint32_t l = d <= 1 ? 0u : 32u - (uint32_t)__clz((int)(d - 1));   // ceil(log2(d))
    if ((1ull << l) == d) {
        f.m = 0;
        f.sh = l;
    } else {
        f.m = (uint32_t)((((1ull << l) - d) << 32) / d) + 1;
        f.sh = l - 1;
    }
    return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv &f) {
    if (f.m == 0) return n >> f.sh;
    uint32_t t = __umulhi(f.m, n);
    return (t + ((n - t) >> 1)) >> f.sh;
}
__device__ __forceinline__ void unravel3f(uint32_t u, int W, int H, const FastDiv &fW, const FastDiv &fH, int &x, int &y, int &z) {
    uint32_t row = fdiv(u, fW);
    x = (int)(u - row * (uint32_t)W);
    uint32_t zz = fdiv(row, fH);
    y = (int)(row - zz * (uint32_t)H);
    z = (int)zz;
}

// (x, y, z) of a raveled index inside a tile / block (all such indices are < 2^31): 32-bit divisions only
__device__ __forceinline__ void unravel3(long long i, int W, int H, int &x, int &y, int &z) {
    uint32_t u = (uint32_t)i;
    uint32_t row = u / (uint32_t)W;
    x = (int)(u - row * (uint32_t)W);
    uint32_t zz = row / (uint32_t)H;
    y = (int)(row - zz * (uint32_t)H);
    z = (int)zz;
}

// fragment id -> dense node index (blocks ascending by id, ids 1..n per block + block_id * prod(block_size)).
// The division by the run-constant prod(block_size) is a multiply-high with a precomputed magic number
// (round-up method, exact for every 64-bit id): q = (((n - mulhi(m, n)) >> 1) + mulhi(m, n)) >> sh.
struct IdMap {
    const uint32_t *cantor2dense;   // block_id - min_block_id -> dense base of that block (0xFFFFFFFF if unknown)
    long long min_block_id, max_block_id;
    long long nvox_block;
    unsigned long long magic;       // 0: nvox_block is a power of two, divide by shifting `sh`
    int sh;
    void set_divisor(long long d) {
        nvox_block = d;
        unsigned long long ud = (unsigned long long)d;
        int l = 0;
        while ((1ull << l) < ud && l < 63) l++;
        if ((1ull << l) == ud) {
            magic = 0;
            sh = l;
        } else {
            unsigned __int128 num = ((unsigned __int128)((1ull << l) - ud)) << 64;
            magic = (unsigned long long)(num / ud) + 1;
            sh = l - 1;
        }
    }
};
__device__ __forceinline__ uint64_t id_block(const IdMap &m, uint64_t id) {
    if (m.magic == 0) return id >> m.sh;
    uint64_t t = __umul64hi(m.magic, id);
    return (((id - t) >> 1) + t) >> m.sh;
}
__device__ __forceinline__ uint32_t id_to_dense(const IdMap &m, uint64_t id) {
    if (id == 0) return 0xFFFFFFFFu;
    uint64_t bid = id_block(m, id);
    if ((long long)bid > m.max_block_id || (long long)bid < m.min_block_id) return 0xFFFFFFFFu;
    uint32_t base = m.cantor2dense[bid - (uint64_t)m.min_block_id];
    if (base == 0xFFFFFFFFu) return 0xFFFFFFFFu;
    return base + (uint32_t)(id - bid * (uint64_t)m.nvox_block) - 1u;
}

// L2-coherent accesses for data exchanged between lanes / CTAs inside one kernel
__device__ __forceinline__ uint32_t ld_cg(const uint32_t *p) { return __ldcg(p); }
__device__ __forceinline__ void st_cg(uint32_t *p, uint32_t v) { __stcg(p, v); }

// lock-free union-find on u32 parents; the root of a set is its minimum index
__device__ __forceinline__ uint32_t uf_find(const uint32_t *parent, uint32_t x) {
    uint32_t p = __ldcg(&parent[x]);
    while (p != x) {
        x = p;
        p = __ldcg(&parent[x]);
    }
    return x;
}
__device__ __forceinline__ void uf_union(uint32_t *parent, uint32_t a, uint32_t b) {
    for (;;) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) {
            uint32_t t = a;
            a = b;
            b = t;
        }
        // a > b: hang a under b if a is still a root
        uint32_t old = atomicMin(&parent[a], b);
        if (old == a) return;
        a = old;
    }
}

}  // namespace bs
