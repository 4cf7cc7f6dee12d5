// Per-label statistics of a label volume (bootstrapper/refine.py:98-108 `_global_sizes`, :236-255 the
// z-extent scan of `z_filter`): voxel count and first / last z plane of every non-zero id, ids ascending.
// One pass with a global open-addressing table (64-bit keys); a warp aggregates its 32 consecutive voxels with
// __match_any_sync, so a run of equal labels costs one insert and three atomics.  The table is compacted, sorted by id
// (LSD radix sort) and gathered.  HBM-bound: 8 bytes read per voxel.
#include "common.cuh"

namespace bs {

static constexpr unsigned long long LS_EMPTY = 0xFFFFFFFFFFFFFFFFull;

__global__ void __launch_bounds__(256) k_label_stats(const uint64_t *__restrict__ seg, size_t n, uint32_t plane, uint32_t tmask,
                                                     unsigned long long *__restrict__ keys, unsigned long long *__restrict__ cnt,
                                                     int *__restrict__ zmin, int *__restrict__ zmax, unsigned int *__restrict__ err) {
    const FastDiv fP = make_fastdiv_dev(plane);
    const int lane = threadIdx.x & 31;
    for (size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) & ~(size_t)31; i0 < n; i0 += (size_t)gridDim.x * blockDim.x) {
        const size_t i = i0 + lane;
        const uint64_t id = i < n ? seg[i] : 0;
        const unsigned act = __ballot_sync(0xFFFFFFFFu, id != 0);
        if (id != 0) {
            const unsigned peers = __match_any_sync(act, (unsigned long long)id);
            const int z = (int)fdiv((uint32_t)i, fP);                         // n < 2^32 (host wrapper)
            const int zhi = __shfl_sync(peers, z, 31 - __clz(peers));         // indices ascend with the lane
            if (lane == __ffs(peers) - 1) {
                uint32_t h = (uint32_t)((id * 0x9E3779B97F4A7C15ull) >> 32) & tmask;
                for (uint32_t probes = 0;; probes++) {
                    const unsigned long long old = atomicCAS(&keys[h], LS_EMPTY, (unsigned long long)id);
                    if (old == LS_EMPTY || old == id) break;
                    if (probes > tmask) {
                        atomicExch(err, 1u);
                        h = 0xFFFFFFFFu;
                        break;
                    }
                    h = (h + 1) & tmask;
                }
                if (h != 0xFFFFFFFFu) {
                    atomicAdd(&cnt[h], (unsigned long long)__popc(peers));
                    atomicMin(&zmin[h], z);
                    atomicMax(&zmax[h], zhi);
                }
            }
        }
    }
}

__global__ void k_ls_flag(const unsigned long long *__restrict__ keys, uint8_t *__restrict__ flag, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = keys[i] != LS_EMPTY ? 1 : 0;
}
__global__ void k_ls_compact(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ pos, size_t n,
                             uint64_t *__restrict__ okeys, uint32_t *__restrict__ oslot) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && keys[i] != LS_EMPTY) {
        okeys[pos[i]] = keys[i];
        oslot[pos[i]] = (uint32_t)i;
    }
}
__global__ void k_ls_gather(const uint64_t *__restrict__ skeys, const uint32_t *__restrict__ sslot, size_t m,
                            const unsigned long long *__restrict__ cnt, const int *__restrict__ zmin, const int *__restrict__ zmax,
                            uint64_t *__restrict__ ids, int64_t *__restrict__ sizes, int32_t *__restrict__ zlo, int32_t *__restrict__ zhi) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) {
        const uint32_t h = sslot[i];
        ids[i] = skeys[i];
        sizes[i] = (int64_t)cnt[h];
        zlo[i] = zmin[h];
        zhi[i] = zmax[h];
    }
}

// seg (Z,Y,X) u64 device; outputs: device arrays of `capacity` entries; n_out host.  Returns BS_ERR_OVERFLOW (with
// *n_out = the number of ids found so far, a lower bound) when the volume holds more than `capacity` distinct ids.
int label_stats(const uint64_t *seg, const int32_t *shape, int64_t capacity, uint64_t *ids, int64_t *sizes, int32_t *zlo,
                int32_t *zhi, int64_t *n_out, cudaStream_t s) {
    BS_ARG(seg && shape && ids && sizes && zlo && zhi && n_out, "bs_label_stats: null argument");
    const int Z = shape[0], Y = shape[1], X = shape[2];
    BS_ARG(Z > 0 && Y > 0 && X > 0 && capacity > 0, "bs_label_stats: empty volume or capacity");
    const size_t n = (size_t)Z * Y * X;
    BS_ARG(n < 0xFFFFFFFFull, "bs_label_stats: at most 2^32 - 1 voxels per call");
    BS_ARG(capacity < (1LL << 30), "bs_label_stats: capacity too large");
    size_t tsize = 1024;
    while (tsize < 2 * (size_t)capacity) tsize <<= 1;
    DevBuf keys, cnt, zmn, zmx, err, flag, pos, tot, ck, cs, ck2, cs2;
    BS_TRY(keys.alloc_fill(8 * tsize, 0xFF, s));
    BS_TRY(cnt.alloc_zero(8 * tsize, s));
    BS_TRY(zmn.alloc_fill(4 * tsize, 0x7F, s));     // 0x7F7F7F7F: larger than any plane index
    BS_TRY(zmx.alloc_fill(4 * tsize, 0xFF, s));     // -1
    BS_TRY(err.alloc_zero(4, s));
    const unsigned grid = (unsigned)std::min<size_t>(cdiv(n, 256), 148 * 16);
    BS_LAUNCH(k_label_stats, grid, 256, 0, s, seg, n, (uint32_t)Y * (uint32_t)X, (uint32_t)(tsize - 1), keys.as<unsigned long long>(),
              cnt.as<unsigned long long>(), zmn.as<int>(), zmx.as<int>(), err.as<unsigned int>());
    BS_TRY(flag.alloc(tsize, s));
    BS_TRY(pos.alloc(4 * tsize, s));
    BS_TRY(tot.alloc_zero(8, s));
    BS_LAUNCH(k_ls_flag, cdiv(tsize, 256), 256, 0, s, keys.as<unsigned long long>(), flag.as<uint8_t>(), tsize);
    BS_TRY(scan_exclusive_u8(flag.as<uint8_t>(), pos.as<uint32_t>(), tsize, tot.as<uint32_t>(), s));
    uint32_t h_tot = 0, h_err = 0;
    BS_CUDA(cudaMemcpyAsync(&h_tot, tot.p, 4, cudaMemcpyDeviceToHost, s));
    BS_CUDA(cudaMemcpyAsync(&h_err, err.p, 4, cudaMemcpyDeviceToHost, s));
    BS_CUDA(cudaStreamSynchronize(s));
    *n_out = h_tot;
    if (h_err || (int64_t)h_tot > capacity) {
        set_error("bs_label_stats: more distinct ids than `capacity`");
        return BS_ERR_OVERFLOW;
    }
    const size_t m = h_tot;
    if (m == 0) return BS_OK;
    BS_TRY(ck.alloc(8 * m, s));
    BS_TRY(cs.alloc(4 * m, s));
    BS_TRY(ck2.alloc(8 * m, s));
    BS_TRY(cs2.alloc(4 * m, s));
    BS_LAUNCH(k_ls_compact, cdiv(tsize, 256), 256, 0, s, keys.as<unsigned long long>(), pos.as<uint32_t>(), tsize, ck.as<uint64_t>(),
              cs.as<uint32_t>());
    BS_TRY(radix_sort_pairs(ck.as<uint64_t>(), cs.as<uint32_t>(), ck2.as<uint64_t>(), cs2.as<uint32_t>(), m, 0, 64, s));
    BS_LAUNCH(k_ls_gather, cdiv(m, 256), 256, 0, s, ck.as<uint64_t>(), cs.as<uint32_t>(), m, cnt.as<unsigned long long>(), zmn.as<int>(),
              zmx.as<int>(), ids, sizes, zlo, zhi);
    BS_CUDA(cudaStreamSynchronize(s));   // the scratch buffers go out of scope
    BS_CUDA(cudaGetLastError());
    return BS_OK;
}

}  // namespace bs
