// Seeded, block-addressable synthetic affinities on the device (test / bench harness, not the timed
// path).  Same float64 arithmetic, operation for operation, as bootstrapper_b200/synth.py (which the
// CPU tests use), so both generators agree bit for bit; compiled with -fmad=false.
#include "common.cuh"

namespace bs {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t mixw(uint64_t x, long long w) { return splitmix64(x ^ ((uint64_t)w * 0x9E3779B97F4A7C15ull)); }
__device__ __forceinline__ double uniform53(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }

__device__ __forceinline__ long long floordiv(long long a, long long b) {
    long long q = a / b;
    if ((a % b != 0) && ((a < 0) != (b < 0))) q--;
    return q;
}

__device__ void cell_of(uint64_t seed, long long z, long long y, long long x, long long &lab, double &b) {
    const long long PZ = 12, PY = 48, PX = 48;
    long long cz = floordiv(z, PZ), cy = floordiv(y, PY), cx = floordiv(x, PX);
    double d1 = INFINITY, d2 = INFINITY;
    lab = 0;
    for (int dz = -1; dz <= 1; dz++)
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                long long iz = cz + dz, iy = cy + dy, ix = cx + dx;
                uint64_t h = mixw(mixw(mixw(seed, iz), iy), ix);
                double sz = 12.0 * ((double)iz + (0.25 + 0.5 * uniform53(mixw(h, 0))));
                double sy = 48.0 * ((double)iy + (0.25 + 0.5 * uniform53(mixw(h, 1))));
                double sx = 48.0 * ((double)ix + (0.25 + 0.5 * uniform53(mixw(h, 2))));
                double ez = 4.0 * ((double)z - sz), ey = (double)y - sy, ex = (double)x - sx;
                double dd = (ez * ez + ey * ey) + ex * ex;
                long long cid = ((iz + 1024) * 4096 + (iy + 1024)) * 4096 + (ix + 1024);
                if (dd < d1) {
                    d2 = d1;
                    lab = cid;
                    d1 = dd;
                } else {
                    d2 = fmin(d2, dd);
                }
            }
    double delta = sqrt(d2) - sqrt(d1);
    double t = delta / 2.5;
    double t2 = t * t;
    b = 1.0 / (1.0 + t2 * t2);
}

template <typename T>
__global__ void __launch_bounds__(256) k_synth(T *__restrict__ out, int SZ, int SY, int SX, int oz, int oy, int ox, uint64_t seed) {
    const size_t n = (size_t)SZ * SY * SX;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int x = (int)(i % SX), y = (int)((i / SX) % SY), z = (int)(i / ((size_t)SX * SY));
        long long gz = oz + z, gy = oy + y, gx = ox + x;
        long long lp, lq;
        double bp, bq;
        cell_of(seed, gz, gy, gx, lp, bp);
        for (int c = 0; c < 3; c++) {
            long long qz = gz - (c == 0), qy = gy - (c == 1), qx = gx - (c == 2);
            double a = 0.0;
            if (qz >= 0 && qy >= 0 && qx >= 0) {
                cell_of(seed, qz, qy, qx, lq, bq);
                double m = fmax(bp, bq);
                a = (lp == lq) ? 1.0 - m : 0.05 * (1.0 - m);
                uint64_t h = mixw(mixw(mixw(mixw(mixw(seed, c), gz), gy), gx), 7);
                double u = uniform53(h);
                a = a + 0.08 * (u - 0.5);
                a = fmin(fmax(a, 0.0), 1.0);
            }
            if (sizeof(T) == 1)
                out[(size_t)c * n + i] = (T)(int)rint(255.0 * a);
            else
                out[(size_t)c * n + i] = (T)a;
        }
    }
}

int synth_affs(void *out, int dtype, const int32_t *shape, const int32_t *offset, uint64_t seed, cudaStream_t s) {
    size_t n = (size_t)shape[0] * shape[1] * shape[2];
    if (n == 0) return BS_OK;
    unsigned grid = (unsigned)std::min<size_t>(cdiv(n, 256), 148 * 64);
    if (dtype == 0)
        BS_LAUNCH((k_synth<uint8_t>), grid, 256, 0, s, (uint8_t *)out, shape[0], shape[1], shape[2], offset[0], offset[1],
                  offset[2], seed);
    else
        BS_LAUNCH((k_synth<float>), grid, 256, 0, s, (float *)out, shape[0], shape[1], shape[2], offset[0], offset[1], offset[2],
                  seed);
    BS_CUDA(cudaGetLastError());
    return BS_OK;
}

}  // namespace bs
