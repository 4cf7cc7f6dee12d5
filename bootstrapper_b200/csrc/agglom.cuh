// Shared declarations of the waterz agglomeration kernels (stage2.cu, agglom_par.cu, agglom_pq.cu).
// Internal header.
#pragma once
#include "geom.h"

namespace bs {

struct AggArrays {
    // edges (global index = ebase[b] + local); eu / ev are block-compact node numbers
    const uint32_t *eu, *ev, *ecnt;
    const unsigned long long *esum;
    // merge tree (per block base 2 * vbase) and merge history (base vbase), in block-compact numbering
    uint32_t *tparent, *tlevel;
    float *tscore;
    uint32_t *ha, *hb;
    float *hs;
    uint32_t *nmerges;
    uint32_t *counters;   // per block: pops, stale, dead, iterations, list-walk batches, append rounds
    uint32_t *error;
};

struct AggBlk {
    uint32_t ebase, E, vbase, nv;   // edge range, compact node range
};

template <bool U8>
__device__ __forceinline__ float edge_score(unsigned long long isum, uint32_t cnt) {
    // OneMinus<MeanAffinity>: (float)(1.0 - mean), mean = float(sum) / float(count)   (oracle edge_score)
    float sum;
    if (U8)
        sum = __double2float_rn(__ddiv_rn((double)isum, 255.0));
    else
        sum = __double2float_rn(ldexp((double)(long long)isum, -38));
    float mean = __fdiv_rn(sum, __uint2float_rn(cnt));
    return __double2float_rn(__dsub_rn(1.0, (double)mean));
}

__device__ __forceinline__ int score_bin(float score, int nbins) {
    int i = (int)__fmul_rn(score, (float)nbins);
    return min(max(0, i), nbins - 1);
}

// agglom_par.cu: the same agglomeration by a CTA of several warps that commits independent merges of a batch together
size_t agglom_par_bytes(uint32_t Ecap, uint32_t Ncap, bool sum64, int idx);
size_t agglom_par_static_smem();
int agglom_par_launch(const AggBlk *blks, const int *list, int nlist, const AggArrays &A, float threshold, int keep_cheaper,
                      bool u8, bool sum64, uint32_t Ecap, uint32_t Ncap, cudaStream_t s);
int agglom_par_global_launch(const AggBlk *blks, const int *list, int nlist, const AggArrays &A, float threshold, int keep_cheaper,
                             bool u8, unsigned char *work, const unsigned long long *woff, uint32_t Ncap_max, bool hybrid,
                             cudaStream_t s);

// agglom_pq.cu: waterz with the non-discretised queue (single-shot ws path) on one region graph.
struct PqRequest {
    const float *thresholds;   // host, ascending
    int T;
    uint64_t *const *segs;     // host array of T device pointers (roi-shaped uint64)
    uint32_t counters[4];      // out: pops, stale, deleted, merges
    int quantile = 0;          // 0: OneMinus<MeanAffinity>; Q: OneMinus<HistogramQuantileAffinity<Q, 256 bins, initmax>>
    int initmax = 0;
};
// hist: nullptr (mean affinity) or [E][256] affinity histograms of the edges (quantile = Q)
int agglom_pq_run(bool u8, uint32_t E, uint32_t Nc, const uint32_t *ceu, const uint32_t *cev, unsigned long long *esum,
                  uint32_t *ecnt, uint32_t *hist, int quantile, const float *thresholds_host, int T, int keep_cheaper,
                  uint32_t *roots_out, uint32_t *counters_host, cudaStream_t s);
int agglom_pq_relabel(const uint64_t *frags, size_t n, IdMap idm, const uint32_t *roots, const uint32_t *cscan, const uint8_t *used,
                      uint32_t Nc, uint32_t nview, uint32_t dense0, long long block_id, long long nvox_block, int T,
                      uint64_t *const *segs, cudaStream_t s);

}  // namespace bs
