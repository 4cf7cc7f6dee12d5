// `bs segment --cc`: connected components of the thresholded short-range affinities.
//
// Replaces cc_affs / compute_connected_component_segmentation (post/connected_components.py:15-127,
// post/cc.py:7-74): hard = affs[:3] > threshold; voxel p is joined with p + e_d when hard[d][p] (the reference's
// "+e_d" convention); a voxel is labelled when one of its own three affinities is set or a lower neighbour's
// affinity points at it; ids are uint32 in raster order of each component's first voxel (the serial flood starts
// a component at its raster-first voxel).  remove_debris = skimage remove_small_objects on the label values.
// Parallel form: lock-free min-root union-find over the voxels, a root bitmap, popcount ranks.
#include <algorithm>

#include "geom.h"

namespace bs {

static constexpr uint32_t NONE32 = 0xFFFFFFFFu;

template <typename T>
__device__ __forceinline__ bool hard_aff(const T *__restrict__ a, const uint8_t *__restrict__ mask, size_t n, int d, size_t i, float thr);
template <>
__device__ __forceinline__ bool hard_aff<uint8_t>(const uint8_t *__restrict__ a, const uint8_t *__restrict__ mask, size_t n, int d,
                                                  size_t i, float thr) {
    // affs.astype(float32) / 255.0, times (mask > 0), compared in float32 (connected_components.py:52-56,66,81)
    float v = __fdiv_rn((float)a[(size_t)d * n + i], 255.0f);
    if (mask && mask[i] == 0) v = 0.0f;
    return v > thr;
}
template <>
__device__ __forceinline__ bool hard_aff<float>(const float *__restrict__ a, const uint8_t *__restrict__ mask, size_t n, int d,
                                                size_t i, float thr) {
    float v = a[(size_t)d * n + i];
    if (mask && mask[i] == 0) v = __fmul_rn(v, 0.0f);
    return v > thr;
}

template <typename T>
__global__ void __launch_bounds__(256) k_ccaff_union(const T *__restrict__ a, const uint8_t *__restrict__ mask, int Z, int Y, int X,
                                                     float thr, uint32_t *__restrict__ parent) {
    const size_t n = (size_t)Z * Y * X;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int x, y, z;
        unravel3((long long)i, X, Y, x, y, z);
        if (z + 1 < Z && hard_aff<T>(a, mask, n, 0, i, thr)) uf_union(parent, (uint32_t)i, (uint32_t)(i + (size_t)Y * X));
        if (y + 1 < Y && hard_aff<T>(a, mask, n, 1, i, thr)) uf_union(parent, (uint32_t)i, (uint32_t)(i + X));
        if (x + 1 < X && hard_aff<T>(a, mask, n, 2, i, thr)) uf_union(parent, (uint32_t)i, (uint32_t)(i + 1));
    }
}

template <typename T>
__device__ __forceinline__ bool cc_labelled(const T *__restrict__ a, const uint8_t *__restrict__ mask, size_t n, int Y, int X, size_t i,
                                            int x, int y, int z, float thr) {
    if (hard_aff<T>(a, mask, n, 0, i, thr) || hard_aff<T>(a, mask, n, 1, i, thr) || hard_aff<T>(a, mask, n, 2, i, thr)) return true;
    if (z > 0 && hard_aff<T>(a, mask, n, 0, i - (size_t)Y * X, thr)) return true;
    if (y > 0 && hard_aff<T>(a, mask, n, 1, i - X, thr)) return true;
    if (x > 0 && hard_aff<T>(a, mask, n, 2, i - 1, thr)) return true;
    return false;
}

template <typename T>
__global__ void __launch_bounds__(256) k_ccaff_roots(const T *__restrict__ a, const uint8_t *__restrict__ mask, int Z, int Y, int X,
                                                     float thr, const uint32_t *__restrict__ parent, uint32_t *__restrict__ bits) {
    const size_t n = (size_t)Z * Y * X;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (parent[i] != (uint32_t)i) continue;
        int x, y, z;
        unravel3((long long)i, X, Y, x, y, z);
        if (cc_labelled<T>(a, mask, n, Y, X, i, x, y, z, thr)) atomicOr(&bits[i >> 5], 1u << (i & 31));
    }
}

__global__ void k_ccaff_popc(const uint32_t *__restrict__ bits, uint32_t *__restrict__ cnt, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cnt[i] = __popc(bits[i]);
}

// label = 1 + rank of the component's root among the roots; a labelled voxel's root is always a labelled root
__global__ void __launch_bounds__(256) k_ccaff_label(size_t n, const uint32_t *__restrict__ parent, const uint32_t *__restrict__ bits,
                                                     const uint32_t *__restrict__ wscan, uint64_t *__restrict__ out,
                                                     uint32_t *__restrict__ sizes) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t r = uf_find(parent, (uint32_t)i);
        uint64_t l = 0;
        if ((bits[r >> 5] >> (r & 31)) & 1u) {
            l = (uint64_t)(wscan[r >> 5] + __popc(bits[r >> 5] & ((1u << (r & 31)) - 1u))) + 1;
            if (sizes) atomicAdd(&sizes[l], 1u);
        }
        out[i] = l;
    }
}

__global__ void __launch_bounds__(256) k_ccaff_debris(size_t n, const uint64_t *__restrict__ lab, const uint32_t *__restrict__ sizes,
                                                      uint32_t min_size, uint64_t *__restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t l = lab[i];
        out[i] = (l && sizes[l] < min_size) ? 0 : l;
    }
}

__global__ void k_ccaff_iota(uint32_t *p, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (uint32_t)i;
}

template <typename T>
static int cc_affs_impl(const T *affs, const uint8_t *mask, int Z, int Y, int X, float thr, int remove_debris, uint64_t *frags_out,
                        uint64_t *seg_out, int64_t *n_out, cudaStream_t s) {
    const size_t n = (size_t)Z * Y * X;
    BS_ARG(n > 0 && n < (1ull << 32) - 1, "bs_cc_affs: volume must have fewer than 2^32 voxels");
    const size_t nw = (n + 31) / 32 + 1;
    DevBuf parent, bits, wcnt, wscan, tot, sizes;
    BS_TRY(parent.alloc(4 * n, s));
    BS_TRY(bits.alloc_zero(4 * nw, s));
    BS_TRY(wcnt.alloc(4 * nw, s));
    BS_TRY(wscan.alloc(4 * nw, s));
    BS_TRY(tot.alloc_zero(16, s));
    const unsigned grid = (unsigned)std::min<size_t>(cdiv(n, 256), 148 * 64);
    BS_LAUNCH(k_ccaff_iota, cdiv(n, 256), 256, 0, s, parent.as<uint32_t>(), n);
    BS_LAUNCH((k_ccaff_union<T>), grid, 256, 0, s, affs, mask, Z, Y, X, thr, parent.as<uint32_t>());
    BS_LAUNCH((k_ccaff_roots<T>), grid, 256, 0, s, affs, mask, Z, Y, X, thr, parent.as<uint32_t>(), bits.as<uint32_t>());
    BS_LAUNCH(k_ccaff_popc, cdiv(nw, 256), 256, 0, s, bits.as<uint32_t>(), wcnt.as<uint32_t>(), nw);
    BS_TRY(scan_exclusive_u32(wcnt.as<uint32_t>(), wscan.as<uint32_t>(), nw, tot.as<uint32_t>(), s));
    uint32_t ncomp = 0;
    BS_CUDA(cudaMemcpyAsync(&ncomp, tot.p, 4, cudaMemcpyDeviceToHost, s));
    BS_CUDA(cudaStreamSynchronize(s));
    if (n_out) *n_out = ncomp;
    const bool debris = remove_debris > 0 && seg_out != nullptr;
    if (debris) BS_TRY(sizes.alloc_zero(4 * ((size_t)ncomp + 2), s));
    BS_LAUNCH(k_ccaff_label, grid, 256, 0, s, n, parent.as<uint32_t>(), bits.as<uint32_t>(), wscan.as<uint32_t>(), frags_out,
              debris ? sizes.as<uint32_t>() : nullptr);
    if (seg_out) {
        if (debris)
            BS_LAUNCH(k_ccaff_debris, grid, 256, 0, s, n, frags_out, sizes.as<uint32_t>(), (uint32_t)remove_debris, seg_out);
        else if (seg_out != frags_out)
            BS_CUDA(cudaMemcpyAsync(seg_out, frags_out, 8 * n, cudaMemcpyDeviceToDevice, s));
    }
    BS_CUDA(cudaStreamSynchronize(s));
    BS_CUDA(cudaGetLastError());
    return BS_OK;
}

int cc_affs(const void *affs, int dtype, const uint8_t *mask, int Z, int Y, int X, float thr, int remove_debris, uint64_t *frags_out,
            uint64_t *seg_out, int64_t *n_out, cudaStream_t s) {
    if (dtype == BS_DTYPE_U8) return cc_affs_impl<uint8_t>((const uint8_t *)affs, mask, Z, Y, X, thr, remove_debris, frags_out, seg_out, n_out, s);
    return cc_affs_impl<float>((const float *)affs, mask, Z, Y, X, thr, remove_debris, frags_out, seg_out, n_out, s);
}

}  // namespace bs
