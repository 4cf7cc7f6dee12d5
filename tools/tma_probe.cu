// Stand-alone probe of the TMA box load used by k_mask_bits_tma (tools only; not part of libbsnative).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu ; ./tma_probe VARIANT
// variants: 0 = 2-D map, in-bounds box; 1 = 4-D map, in-bounds; 2 = 4-D, negative start coordinates (zero fill);
//           3 = 4-D, descriptor passed as __grid_constant__ parameter; 4 = 4-D box reaching past the end of x and y
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

typedef CUresult (*enc_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                           const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(phase) : "memory");
    }
}

template <int RANK>
__global__ void k_probe(const CUtensorMap *gmap, const __grid_constant__ CUtensorMap pmap, int use_param, int x, int y, int z, int c,
                        uint8_t *out, int box_bytes) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const CUtensorMap *m = use_param ? &pmap : gmap;
        mbar_expect_tx(&bar, (uint32_t)box_bytes);
        if (RANK == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                             (uint32_t)__cvta_generic_to_shared(sm)),
                         "l"((uint64_t)m), "r"((uint32_t)__cvta_generic_to_shared(&bar)), "r"(x), "r"(y)
                         : "memory");
        else
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                             (uint32_t)__cvta_generic_to_shared(sm)),
                         "l"((uint64_t)m), "r"((uint32_t)__cvta_generic_to_shared(&bar)), "r"(x), "r"(y), "r"(z), "r"(c)
                         : "memory");
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < box_bytes; i += blockDim.x) out[i] = sm[i];
}

int main(int argc, char **argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int C = 3, Z = 4, Y = 64, X = 160, BW = 64, BH = 16;
    std::vector<uint8_t> h((size_t)C * Z * Y * X);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 7 + (i >> 8));
    uint8_t *d, *out;
    cudaMalloc(&d, h.size());
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    cudaMalloc(&out, BW * BH);
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
        printf("variant %d: no cuTensorMapEncodeTiled\n", variant);
        return 2;
    }
    enc_fn enc = (enc_fn)sym;
    CUtensorMap map;
    CUresult r;
    const int rank = variant == 0 ? 2 : 4;
    if (rank == 2) {
        const cuuint64_t gd[2] = {(cuuint64_t)X, (cuuint64_t)Y * Z * C};
        const cuuint64_t gs[1] = {(cuuint64_t)X};
        const cuuint32_t box[2] = {BW, BH}, es[2] = {1, 1};
        r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        const cuuint64_t gd[4] = {(cuuint64_t)X, (cuuint64_t)Y, (cuuint64_t)Z, (cuuint64_t)C};
        const cuuint64_t gs[3] = {(cuuint64_t)X, (cuuint64_t)X * Y, (cuuint64_t)X * Y * Z};
        const cuuint32_t box[4] = {BW, BH, 1, 1}, es[4] = {1, 1, 1, 1};
        r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) {
        printf("variant %d: encode failed %d\n", variant, (int)r);
        return 3;
    }
    CUtensorMap *dmap;
    cudaMalloc(&dmap, sizeof(CUtensorMap));
    cudaMemcpy(dmap, &map, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    int x = 16, y = 8, z = 1, c = 2;
    if (variant == 2) x = -8, y = -4;
    if (variant == 4) x = X - 32, y = Y - 8;
    if (variant == 5) x = 8;
    if (variant == 6) x = -16;
    if (variant == 7) y = -4;
    if (variant == 8) x = -8;
    if (variant == 9) x = 3;
    if (rank == 2)
        k_probe<2><<<1, 128, BW * BH>>>(dmap, map, 0, x, y, 0, 0, out, BW * BH);
    else
        k_probe<4><<<1, 128, BW * BH>>>(dmap, map, variant == 3, x, y, z, c, out, BW * BH);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("variant %d: kernel failed: %s\n", variant, cudaGetErrorString(e));
        return 1;
    }
    std::vector<uint8_t> o(BW * BH);
    cudaMemcpy(o.data(), out, o.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int j = 0; j < BH; j++)
        for (int i = 0; i < BW; i++) {
            const int xx = x + i, yy = y + j;
            uint8_t want = 0;
            if (rank == 2) {
                if (xx >= 0 && xx < X && yy >= 0 && yy < Y * Z * C) want = h[(size_t)yy * X + xx];
            } else {
                if (xx >= 0 && xx < X && yy >= 0 && yy < Y) want = h[(((size_t)c * Z + z) * Y + yy) * X + xx];
            }
            bad += o[j * BW + i] != want;
        }
    printf("variant %d: ok, %d mismatching bytes\n", variant, bad);
    return bad ? 4 : 0;
}
