// scipy.ndimage.gaussian_filter(affs, sigma=(0, sz, sy, sx)) replayed bit for bit -- the `sigma` shift of the ws / cc
// paths (watershed_frags.py:121-122, post/watershed.py:291-293, connected_components.py:73-75).
//
// scipy runs one correlate1d pass per axis with sigma > 1e-15, in axis order, each pass computing in double
//     tmp = line[l] * w[r];  for ii = -r .. -1:  tmp += (line[l + ii] + line[l - ii]) * w[ii + r]
// on the 'reflect'-extended line and storing in the array's dtype (float64 for the uint8-derived blockwise input,
// float32 for float32 input) before the next pass.  The weights come from the host (numpy: exp, normalise), this file
// only replays the sums; the library is compiled with -fmad=false, as scipy's loop is not contracted either (pinned
// against scipy itself in tests/test_gpu_parity.py).
#include "geom.h"

namespace bs {

__device__ __forceinline__ int g_reflect(int j, int L) {
    while (j < 0 || j >= L) {
        if (j < 0) j = -j - 1;
        if (j >= L) j = 2 * L - j - 1;
    }
    return j;
}

// normalised, masked, zero-filled affinities of every tile (3 channels), in the dtype numpy holds them in
template <typename T, typename Out>
__global__ void __launch_bounds__(256) k_gauss_load(const Tile *__restrict__ tiles, const T *__restrict__ a, const uint8_t *__restrict__ mask,
                                                    int volZw, int volZ, int volY, int volX, int z0, size_t cstride,
                                                    Out *__restrict__ out) {
    const Tile t = tiles[blockIdx.y];
    const long long npix = (long long)t.D * t.H * t.W;
    const size_t nvol = (size_t)volZw * volY * volX;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
        int x, y, z;
        unravel3(i, t.W, t.H, x, y, z);
        const int gz = t.gz + z, gy = t.gy + y, gx = t.gx + x;
        const bool inside = gz >= 0 && gz < volZ && gz >= z0 && gz < z0 + volZw && gy >= 0 && gy < volY && gx >= 0 && gx < volX;
        const size_t gi = inside ? ((size_t)(gz - z0) * volY + gy) * volX + gx : 0;
        const bool on = inside && (!mask || mask[gi] > 0);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            Out v = 0;
            if (on) {
                if constexpr (sizeof(T) == 1) {
                    if constexpr (sizeof(Out) == 8)
                        v = (Out)__ddiv_rn((double)a[(size_t)c * nvol + gi], 255.0);        // astype(float64) / 255
                    else
                        v = (Out)__fdiv_rn((float)a[(size_t)c * nvol + gi], 255.0f);        // astype(float32) / 255
                } else {
                    v = (Out)a[(size_t)c * nvol + gi];
                }
            }
            out[(size_t)c * cstride + t.base + i] = v;
        }
    }
}

template <typename Out>
__global__ void __launch_bounds__(256) k_gauss_axis(const Tile *__restrict__ tiles, int axis, int radius, const double *__restrict__ w,
                                                    size_t cstride, const Out *__restrict__ in, Out *__restrict__ out) {
    const Tile t = tiles[blockIdx.y];
    const int W = t.W, H = t.H, D = t.D;
    const long long HW = (long long)H * W, npix = (long long)D * HW;
    const Out *ip = in + (size_t)blockIdx.z * cstride + t.base;
    Out *op = out + (size_t)blockIdx.z * cstride + t.base;
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
        const Out *line = ip + i - (long long)c * st;
        double tmp = __dmul_rn((double)line[(long long)c * st], w[radius]);
        for (int ii = -radius; ii < 0; ii++) {
            const double a = (double)line[(long long)g_reflect(c + ii, L) * st], b = (double)line[(long long)g_reflect(c - ii, L) * st];
            tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(a, b), w[ii + radius]));
        }
        op[i] = (Out)tmp;
    }
}

// out (3 channels, tile layout) = gaussian_filter of the tiles' affinities; `scratch` is a second buffer of the same size.
// Returns which of the two buffers holds the result.
template <typename T, typename Out>
int gauss_tiles(const Tile *d_tiles, int ntiles, long long maxpix, const T *affs, const uint8_t *mask, int volZw, int volZ, int volY,
                int volX, int z0, const GaussWeights &gw, size_t cstride, Out *bufA, Out *bufB, Out **result, cudaStream_t s) {
    const dim3 grid((unsigned)std::min<long long>(std::max<long long>((maxpix + 4095) / 4096, 1), 2048), ntiles);
    BS_LAUNCH((k_gauss_load<T, Out>), grid, 256, 0, s, d_tiles, affs, mask, volZw, volZ, volY, volX, z0, cstride, bufA);
    Out *cur = bufA, *nxt = bufB;
    for (int axis = 0; axis < 3; axis++) {
        if (gw.radius[axis] < 0) continue;
        const dim3 g3(grid.x, ntiles, 3);
        BS_LAUNCH((k_gauss_axis<Out>), g3, 256, 0, s, d_tiles, axis, gw.radius[axis], gw.w[axis], cstride, cur, nxt);
        std::swap(cur, nxt);
    }
    *result = cur;
    BS_CUDA(cudaGetLastError());
    return BS_OK;
}

template int gauss_tiles<uint8_t, double>(const Tile *, int, long long, const uint8_t *, const uint8_t *, int, int, int, int, int,
                                          const GaussWeights &, size_t, double *, double *, double **, cudaStream_t);
template int gauss_tiles<float, float>(const Tile *, int, long long, const float *, const uint8_t *, int, int, int, int, int,
                                       const GaussWeights &, size_t, float *, float *, float **, cudaStream_t);
template int gauss_tiles<uint8_t, float>(const Tile *, int, long long, const uint8_t *, const uint8_t *, int, int, int, int, int,
                                         const GaussWeights &, size_t, float *, float *, float **, cudaStream_t);

// ---- single-shot paths: out (3, Z, Y, X) float32 = affs_data + shift exactly as numpy computes it
//   affs_data = affs[:3].astype(float32) (/ 255 for uint8), *= (mask > 0);  shift = zeros_like;
//   shift += gaussian_filter(affs_data, (0, *sigma)) - affs_data;  shift += bias (a float64 array);  affs_data += shift
// (post/watershed.py:262-303, connected_components.py:52-77)
__global__ void __launch_bounds__(256) k_shift_combine(size_t n, const float *__restrict__ x, const float *__restrict__ g, int has_sigma,
                                                       int has_bias, double b0, double b1, double b2, float *__restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * n) return;
    const int c = (int)(i / n);
    const float xv = x[i];
    float sh = 0.0f;
    if (has_sigma) sh = __fadd_rn(sh, __fsub_rn(g[i], xv));
    if (has_bias) sh = __double2float_rn(__dadd_rn((double)sh, c == 0 ? b0 : (c == 1 ? b1 : b2)));
    out[i] = __fadd_rn(xv, sh);
}

int shift_affinities(const void *affs, int dtype, const uint8_t *mask, int Z, int Y, int X, int has_sigma, const int *radius,
                     const double *const *w_host, int has_bias, const double *bias, float *out, cudaStream_t s) {
    const size_t n = (size_t)Z * Y * X;
    BS_ARG(n > 0 && n < (1ull << 31), "bs_shift_affinities: volume must have fewer than 2^31 voxels");
    Tile t;
    t.gz = t.gy = t.gx = 0;
    t.D = Z, t.H = Y, t.W = X;
    t.wz = t.wy = t.wx = 0;
    t.wD = Z, t.wH = Y, t.wW = X;
    t.block = 0, t.ndim = 3, t.base = 0, t.wbase = 0;
    t.set_divs();
    DevBuf d_t, A, B, dw[3];
    BS_TRY(d_t.alloc(sizeof(Tile), s));
    BS_CUDA(cudaMemcpyAsync(d_t.p, &t, sizeof(Tile), cudaMemcpyHostToDevice, s));
    BS_TRY(A.alloc(12 * n, s));
    BS_TRY(B.alloc(12 * n, s));
    GaussWeights gw;
    for (int ax = 0; ax < 3; ax++) {
        gw.radius[ax] = has_sigma ? radius[ax] : -1;
        gw.w[ax] = nullptr;
        if (gw.radius[ax] >= 0) {
            BS_TRY(dw[ax].alloc(8 * (size_t)(2 * radius[ax] + 1), s));
            BS_CUDA(cudaMemcpyAsync(dw[ax].p, w_host[ax], 8 * (size_t)(2 * radius[ax] + 1), cudaMemcpyHostToDevice, s));
            gw.w[ax] = dw[ax].as<double>();
        }
    }
    BS_CUDA(cudaStreamSynchronize(s));   // host-staged copies
    float *res = nullptr;
    if (dtype == BS_DTYPE_U8)
        BS_TRY((gauss_tiles<uint8_t, float>(d_t.as<Tile>(), 1, (long long)n, (const uint8_t *)affs, mask, Z, Z, Y, X, 0, gw, n, A.as<float>(),
                                            B.as<float>(), &res, s)));
    else
        BS_TRY((gauss_tiles<float, float>(d_t.as<Tile>(), 1, (long long)n, (const float *)affs, mask, Z, Z, Y, X, 0, gw, n, A.as<float>(),
                                          B.as<float>(), &res, s)));
    // x = the unfiltered normalised affinities: reload them into the other buffer
    float *xbuf = res == A.as<float>() ? B.as<float>() : A.as<float>();
    GaussWeights none;
    for (int ax = 0; ax < 3; ax++) none.radius[ax] = -1, none.w[ax] = nullptr;
    float *xres = nullptr;
    DevBuf dummy;
    BS_TRY(dummy.alloc(16, s));
    if (dtype == BS_DTYPE_U8)
        BS_TRY((gauss_tiles<uint8_t, float>(d_t.as<Tile>(), 1, (long long)n, (const uint8_t *)affs, mask, Z, Z, Y, X, 0, none, n, xbuf,
                                            dummy.as<float>(), &xres, s)));
    else
        BS_TRY((gauss_tiles<float, float>(d_t.as<Tile>(), 1, (long long)n, (const float *)affs, mask, Z, Z, Y, X, 0, none, n, xbuf,
                                          dummy.as<float>(), &xres, s)));
    BS_LAUNCH(k_shift_combine, cdiv(3 * n, 256), 256, 0, s, n, xbuf, res, has_sigma, has_bias, has_bias ? bias[0] : 0.0,
              has_bias ? bias[1] : 0.0, has_bias ? bias[2] : 0.0, out);
    BS_CUDA(cudaStreamSynchronize(s));
    BS_CUDA(cudaGetLastError());
    return BS_OK;
}

}  // namespace bs
