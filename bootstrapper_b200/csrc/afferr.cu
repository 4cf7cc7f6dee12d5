// Affinity self-consistency error (bootstrapper/gp/add_aff_errors.py:128-183): segmentation -> affinities
// on a neighbourhood, squared difference to the predicted affinities summed over the channels (float32, channel order),
// optional mask, division by the maximum, threshold mask.  Two streaming passes: the map with a block-reduced
// maximum, then the normalisation.  HBM-bound: 8 (seg) + C * sizeof(pred) read, 4 + 4 + 1 (+ 4 C) written per voxel.
#include "common.cuh"

namespace bs {

static constexpr int AE_MAXC = 32;
struct AeNhood {
    int off[AE_MAXC][3];
    long long lin[AE_MAXC];
    int C;
};

template <typename T>
__device__ __forceinline__ float ae_pred(const T *p, size_t i);
template <>
__device__ __forceinline__ float ae_pred<float>(const float *p, size_t i) {
    return p[i];
}
template <>
__device__ __forceinline__ float ae_pred<uint8_t>(const uint8_t *p, size_t i) {
    return __fmul_rn((float)p[i], 1.0f / 255.0f);   // gp.Normalize: astype(float32) * float32(1/255)
}

template <typename T>
__global__ void __launch_bounds__(256) k_afferr_map(const uint64_t *__restrict__ seg, const T *__restrict__ pred,
                                                    const uint8_t *__restrict__ mask, AeNhood nh, int Z, int Y, int X,
                                                    float *__restrict__ seg_affs, float *__restrict__ err,
                                                    unsigned int *__restrict__ maxbits) {
    const size_t n = (size_t)Z * Y * X;
    const FastDiv fX = make_fastdiv_dev((uint32_t)X), fY = make_fastdiv_dev((uint32_t)Y);
    float mymax = 0.f;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x; i0 < n; i0 += (size_t)gridDim.x * blockDim.x) {
        const size_t i = i0 + threadIdx.x;
        if (i < n) {
            // n < 2^32 per call is guaranteed by the host wrapper (chunks of whole z planes)
            int x, y, z;
            unravel3f((uint32_t)i, X, Y, fX, fY, x, y, z);
            const uint64_t s = seg[i];
            float acc = 0.f;
#pragma unroll 1
            for (int c = 0; c < nh.C; c++) {
                const int zz = z + nh.off[c][0], yy = y + nh.off[c][1], xx = x + nh.off[c][2];
                float a = 0.f;
                if (s != 0 && zz >= 0 && zz < Z && yy >= 0 && yy < Y && xx >= 0 && xx < X) a = seg[(long long)i + nh.lin[c]] == s ? 1.f : 0.f;
                if (seg_affs) seg_affs[(size_t)c * n + i] = a;
                const float d = __fsub_rn(a, ae_pred<T>(pred, (size_t)c * n + i));
                const float q = __fmul_rn(d, d);
                acc = c == 0 ? q : __fadd_rn(acc, q);
            }
            if (mask) acc = __fmul_rn(acc, (float)mask[i]);
            err[i] = acc;
            mymax = fmaxf(mymax, acc);
        }
    }
    mymax = __uint_as_float(__reduce_max_sync(0xFFFFFFFFu, __float_as_uint(mymax)));   // non-negative floats order like their bits
    if ((threadIdx.x & 31) == 0 && mymax > 0.f) atomicMax(maxbits, __float_as_uint(mymax));
}

// four consecutive voxels per thread (n % 4 == 0): 32-byte segmentation loads, 16-byte / 4-byte prediction loads, 16-byte
// stores; the neighbour ids come through L1
template <typename T>
struct AePred4;
template <>
struct AePred4<float> {
    static __device__ __forceinline__ void load(const float *p, size_t i, float (&v)[4]) {
        const float4 q = *(const float4 *)(p + i);
        v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
    }
};
template <>
struct AePred4<uint8_t> {
    static __device__ __forceinline__ void load(const uint8_t *p, size_t i, float (&v)[4]) {
        const uchar4 q = *(const uchar4 *)(p + i);
        v[0] = __fmul_rn((float)q.x, 1.0f / 255.0f), v[1] = __fmul_rn((float)q.y, 1.0f / 255.0f);
        v[2] = __fmul_rn((float)q.z, 1.0f / 255.0f), v[3] = __fmul_rn((float)q.w, 1.0f / 255.0f);
    }
};

template <typename T>
__global__ void __launch_bounds__(256) k_afferr_map4(const uint64_t *__restrict__ seg, const T *__restrict__ pred,
                                                     const uint8_t *__restrict__ mask, AeNhood nh, int Z, int Y, int X,
                                                     float *__restrict__ seg_affs, float *__restrict__ err,
                                                     unsigned int *__restrict__ maxbits) {
    const size_t n = (size_t)Z * Y * X, n4 = n >> 2;
    const FastDiv fX = make_fastdiv_dev((uint32_t)X), fY = make_fastdiv_dev((uint32_t)Y);
    float mymax = 0.f;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n4; j += (size_t)gridDim.x * blockDim.x) {
        const size_t i = j << 2;
        int x[4], y[4], z[4];
        unravel3f((uint32_t)i, X, Y, fX, fY, x[0], y[0], z[0]);
#pragma unroll
        for (int k = 1; k < 4; k++) {
            x[k] = x[k - 1] + 1, y[k] = y[k - 1], z[k] = z[k - 1];
            if (x[k] == X) {
                x[k] = 0;
                if (++y[k] == Y) y[k] = 0, z[k]++;
            }
        }
        uint64_t s[4];
        {
            const ulonglong2 a = *(const ulonglong2 *)(seg + i), b = *(const ulonglong2 *)(seg + i + 2);
            s[0] = a.x, s[1] = a.y, s[2] = b.x, s[3] = b.y;
        }
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int c = 0; c < nh.C; c++) {
            float pv[4], a[4];
            AePred4<T>::load(pred, (size_t)c * n + i, pv);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int zz = z[k] + nh.off[c][0], yy = y[k] + nh.off[c][1], xx = x[k] + nh.off[c][2];
                a[k] = 0.f;
                if (s[k] != 0 && zz >= 0 && zz < Z && yy >= 0 && yy < Y && xx >= 0 && xx < X)
                    a[k] = seg[(long long)(i + k) + nh.lin[c]] == s[k] ? 1.f : 0.f;
                const float d = __fsub_rn(a[k], pv[k]);
                const float q = __fmul_rn(d, d);
                acc[k] = c == 0 ? q : __fadd_rn(acc[k], q);
            }
            if (seg_affs) *(float4 *)(seg_affs + (size_t)c * n + i) = make_float4(a[0], a[1], a[2], a[3]);
        }
        if (mask) {
            const uchar4 m = *(const uchar4 *)(mask + i);
            acc[0] = __fmul_rn(acc[0], (float)m.x), acc[1] = __fmul_rn(acc[1], (float)m.y);
            acc[2] = __fmul_rn(acc[2], (float)m.z), acc[3] = __fmul_rn(acc[3], (float)m.w);
        }
        *(float4 *)(err + i) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        mymax = fmaxf(fmaxf(mymax, fmaxf(acc[0], acc[1])), fmaxf(acc[2], acc[3]));
    }
    mymax = __uint_as_float(__reduce_max_sync(0xFFFFFFFFu, __float_as_uint(mymax)));
    if ((threadIdx.x & 31) == 0 && mymax > 0.f) atomicMax(maxbits, __float_as_uint(mymax));
}

__global__ void __launch_bounds__(256) k_afferr_norm4(float *__restrict__ err, uint8_t *__restrict__ emask, size_t n4,
                                                      const unsigned int *__restrict__ maxbits, float floor_, float ceil_) {
    const float mx = __uint_as_float(*maxbits);
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n4; j += (size_t)gridDim.x * blockDim.x) {
        float4 v = ((const float4 *)err)[j];
        v.x = mx > 0.f ? __fdiv_rn(v.x, mx) : 0.f, v.y = mx > 0.f ? __fdiv_rn(v.y, mx) : 0.f;
        v.z = mx > 0.f ? __fdiv_rn(v.z, mx) : 0.f, v.w = mx > 0.f ? __fdiv_rn(v.w, mx) : 0.f;
        ((float4 *)err)[j] = v;
        ((uchar4 *)emask)[j] = make_uchar4((v.x > floor_ && v.x < ceil_) ? 1 : 0, (v.y > floor_ && v.y < ceil_) ? 1 : 0,
                                           (v.z > floor_ && v.z < ceil_) ? 1 : 0, (v.w > floor_ && v.w < ceil_) ? 1 : 0);
    }
}

__global__ void __launch_bounds__(256) k_afferr_norm(float *__restrict__ err, uint8_t *__restrict__ emask, size_t n,
                                                     const unsigned int *__restrict__ maxbits, float floor_, float ceil_) {
    const float mx = __uint_as_float(*maxbits);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = mx > 0.f ? __fdiv_rn(err[i], mx) : 0.f;
        err[i] = v;
        emask[i] = (v > floor_ && v < ceil_) ? 1 : 0;
    }
}

// seg (Z,Y,X) u64, pred (C,Z,Y,X) float32 (dtype 1) or uint8 (dtype 0), nhood (C,3) host ints, mask (Z,Y,X) u8 or null,
// seg_affs (C,Z,Y,X) float32 or null, err (Z,Y,X) float32, emask (Z,Y,X) u8; scratch: 4 bytes (device)
int aff_errors(const uint64_t *seg, const void *pred, int pred_dtype, int C, const int32_t *shape, const int32_t *nhood,
               const uint8_t *mask, float floor_, float ceil_, float *seg_affs, float *err, uint8_t *emask, cudaStream_t s) {
    BS_ARG(seg && pred && shape && nhood && err && emask, "bs_aff_errors: null argument");
    BS_ARG(C >= 1 && C <= AE_MAXC, "bs_aff_errors: 1..32 neighbourhood offsets");
    BS_ARG(pred_dtype == 0 || pred_dtype == 1, "bs_aff_errors: predicted affinities must be uint8 or float32");
    const int Z = shape[0], Y = shape[1], X = shape[2];
    BS_ARG(Z > 0 && Y > 0 && X > 0, "bs_aff_errors: empty volume");
    const size_t n = (size_t)Z * Y * X;
    BS_ARG(n < 0xFFFFFFFFull, "bs_aff_errors: at most 2^32 - 1 voxels per call");
    AeNhood nh;
    nh.C = C;
    for (int c = 0; c < C; c++) {
        for (int d = 0; d < 3; d++) nh.off[c][d] = nhood[3 * c + d];
        nh.lin[c] = ((long long)nhood[3 * c] * Y + nhood[3 * c + 1]) * X + nhood[3 * c + 2];
    }
    DevBuf mx;
    BS_TRY(mx.alloc_zero(4, s));
    const bool vec = (n & 3) == 0 && (((uintptr_t)seg | (uintptr_t)err | (uintptr_t)seg_affs) & 31) == 0 &&
                     (((uintptr_t)pred | (uintptr_t)emask | (uintptr_t)mask) & 15) == 0;
    if (vec) {
        const unsigned grid4 = (unsigned)std::min<size_t>(cdiv(n >> 2, 256), 148 * 32);
        if (pred_dtype == 1)
            BS_LAUNCH(k_afferr_map4<float>, grid4, 256, 0, s, seg, (const float *)pred, mask, nh, Z, Y, X, seg_affs, err, mx.as<unsigned int>());
        else
            BS_LAUNCH(k_afferr_map4<uint8_t>, grid4, 256, 0, s, seg, (const uint8_t *)pred, mask, nh, Z, Y, X, seg_affs, err, mx.as<unsigned int>());
        BS_LAUNCH(k_afferr_norm4, grid4, 256, 0, s, err, emask, n >> 2, mx.as<unsigned int>(), floor_, ceil_);
        BS_CUDA(cudaGetLastError());
        return BS_OK;
    }
    const unsigned grid = (unsigned)std::min<size_t>(cdiv(n, 256), 148 * 32);
    if (pred_dtype == 1)
        BS_LAUNCH(k_afferr_map<float>, grid, 256, 0, s, seg, (const float *)pred, mask, nh, Z, Y, X, seg_affs, err, mx.as<unsigned int>());
    else
        BS_LAUNCH(k_afferr_map<uint8_t>, grid, 256, 0, s, seg, (const uint8_t *)pred, mask, nh, Z, Y, X, seg_affs, err, mx.as<unsigned int>());
    BS_LAUNCH(k_afferr_norm, grid, 256, 0, s, err, emask, n, mx.as<unsigned int>(), floor_, ceil_);
    BS_CUDA(cudaGetLastError());
    return BS_OK;
}

}  // namespace bs
