// extern "C" surface of libbsnative (include/bsnative.h).  No torch types, no CPU fallback.
#include <string.h>

#include <algorithm>

#include <atomic>
#include <thread>
#include <vector>

#include "agglom.cuh"

namespace bs {
int g_debug = 0;
int g_agglom_version = 0;  // 0 = auto (shared-memory kernel when the block fits), 3 / 4 = force the global-slab kernels
int g_front_version = 0;   // stage-1 front end: 0 = auto (fused when eligible), 1 = unfused chain, 2 = fused without TMA, 3 = fused, TMA required
int g_flood_version = 0;   // 0 = auto (v2 when eligible), 1 = force the global-memory flood
int stage1_run(Plan &P, const void *affs, const uint8_t *mask, uint64_t *frags_out, cudaStream_t s, const uint32_t *ext_labels = nullptr,
               size_t ext_nlabels = 0);
int label_stats(const uint64_t *seg, const int32_t *shape, int64_t capacity, uint64_t *ids, int64_t *sizes, int32_t *zlo,
                int32_t *zhi, int64_t *n_out, cudaStream_t s);
int aff_errors(const uint64_t *seg, const void *pred, int pred_dtype, int C, const int32_t *shape, const int32_t *nhood,
               const uint8_t *mask, float floor_, float ceil_, float *seg_affs, float *err, uint8_t *emask, cudaStream_t s);
int stage2_run(Plan &P, const void *affs, const uint64_t *frags, cudaStream_t s, PqRequest *pq);
int connected_components(const uint64_t *nodes, int64_t n, const uint64_t *eu, const uint64_t *ev, const float *scores,
                         int64_t m, float thr, uint64_t *comp, cudaStream_t s);
int relabel(const uint64_t *frags, int64_t n, const uint64_t *keys, const uint64_t *vals, int64_t k, uint64_t *seg,
            cudaStream_t s);
int relabel_dense(Plan &P, const uint64_t *frags, int64_t n, const uint64_t *const *comps, int T, uint64_t *const *segs,
                  cudaStream_t s);
int components_multi(Plan &P, const uint64_t *nodes, int64_t n, const uint64_t *eu, const uint64_t *ev, const float *scores,
                     int64_t m, const float *thresholds, int T, uint64_t *const *comps, cudaStream_t s);
int plan_node_ids(Plan &P, uint64_t *out, long long *n_out, cudaStream_t s);
int aff_agglom(Plan &P, const void *affs, const uint64_t *frags, int C, const int32_t *offsets, cudaStream_t s);
int dense_fragments(Plan &P, const uint64_t *frags, int64_t n, uint32_t *dense_out, cudaStream_t s);
int cc_affs(const void *affs, int dtype, const uint8_t *mask, int Z, int Y, int X, float thr, int remove_debris, uint64_t *frags_out,
            uint64_t *seg_out, int64_t *n_out, cudaStream_t s);
int shift_affinities(const void *affs, int dtype, const uint8_t *mask, int Z, int Y, int X, int has_sigma, const int *radius,
                     const double *const *w_host, int has_bias, const double *bias, float *out, cudaStream_t s);
int synth_affs(void *out, int dtype, const int32_t *shape, const int32_t *offset, uint64_t seed, cudaStream_t s);
}  // namespace bs

using namespace bs;

struct bs_plan {
    Plan *p;
};

static int copy_out(void *dst, const void *src, size_t bytes, cudaStream_t s) {
    if (bytes == 0) return BS_OK;
    BS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, s));
    BS_CUDA(cudaStreamSynchronize(s));
    return BS_OK;
}

extern "C" {

const char *bs_last_error(void) { return get_error(); }
unsigned long long bs_launch_count(void) { return g_launches; }
int bs_version(void) { return 102; }
unsigned long long bs_config_size(void) { return sizeof(bs_ws_config); }
int bs_set_debug(int on) {
    g_debug = on;
    return BS_OK;
}

// keep freed scratch in the stream-ordered pool instead of returning it to the driver at every sync
static void init_mempool() {
    static std::atomic<unsigned long long> done(0);   // one bit per device
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
    if (done.load() & (1ull << dev)) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    done.fetch_or(1ull << dev);
}

int bs_set_flood_version(int v) {
    g_flood_version = v;
    return BS_OK;
}

int bs_set_front_version(int v) {
    g_front_version = v;
    return BS_OK;
}

int bs_set_agglom_version(int v) {
    BS_ARG(v == 0 || v == 3 || v == 4, "bs_set_agglom_version: 0 (automatic), 3 or 4 (the single-warp kernels 1 / 2 were removed)");
    g_agglom_version = v;
    return BS_OK;
}

int bs_plan_create(const bs_ws_config *cfg, bs_plan **out) {
    BS_ARG(cfg && out, "bs_plan_create: null argument");
    init_mempool();
    for (int d = 0; d < 3; d++)
        BS_ARG(cfg->context[d] <= cfg->block_size[d], "bs_plan_create: context larger than the block is not supported");
    Plan *p = nullptr;
    BS_TRY(plan_build(*cfg, &p));
    bs_plan *h = new bs_plan();
    h->p = p;
    *out = h;
    return BS_OK;
}

void bs_plan_destroy(bs_plan *p) {
    if (!p) return;
    delete p->p;
    delete p;
}

int bs_plan_num_blocks(const bs_plan *p, int64_t *n_total, int64_t *n_owned) {
    BS_ARG(p, "null plan");
    if (n_total) *n_total = (int64_t)p->p->blocks.size();
    if (n_owned) *n_owned = (int64_t)p->p->owned.size();
    return BS_OK;
}

int bs_plan_block_info(const bs_plan *p, int64_t *block_id, int32_t *write_offset, int32_t *write_shape) {
    BS_ARG(p, "null plan");
    const auto &bl = p->p->blocks;
    for (size_t i = 0; i < bl.size(); i++) {
        if (block_id) block_id[i] = bl[i].block_id;
        for (int d = 0; d < 3; d++) {
            if (write_offset) write_offset[3 * i + d] = bl[i].wo[d];
            if (write_shape) write_shape[3 * i + d] = bl[i].ws[d];
        }
    }
    return BS_OK;
}

int bs_plan_set_owned(bs_plan *p, const int32_t *indices, int64_t n) {
    BS_ARG(p && (n == 0 || indices), "bs_plan_set_owned: null argument");
    Plan &P = *p->p;
    std::vector<int> own;
    for (auto &b : P.blocks) b.owned = 0;
    for (int64_t i = 0; i < n; i++) {
        BS_ARG(indices[i] >= 0 && (size_t)indices[i] < P.blocks.size(), "bs_plan_set_owned: block index out of range");
        P.blocks[indices[i]].owned = 1;
    }
    for (size_t i = 0; i < P.blocks.size(); i++)
        if (P.blocks[i].owned) own.push_back((int)i);
    P.owned = own;
    return BS_OK;
}

int bs_stage1_fragments(bs_plan *p, const void *affs, const uint8_t *mask, uint64_t *frags_out, void *stream) {
    BS_ARG(p && affs && frags_out, "bs_stage1_fragments: null argument");
    return stage1_run(*p->p, affs, mask, frags_out, (cudaStream_t)stream);
}

int bs_stage1_num_nodes(const bs_plan *p, int64_t *n) {
    BS_ARG(p && n, "null argument");
    *n = p->p->n_nodes;
    return BS_OK;
}

int bs_stage1_get_nodes(const bs_plan *p, uint64_t *ids, int32_t *pos_zyx, uint32_t *sizes, void *stream) {
    BS_ARG(p, "null plan");
    cudaStream_t s = (cudaStream_t)stream;
    size_t n = (size_t)p->p->n_nodes;
    if (ids) BS_TRY(copy_out(ids, p->p->node_id.p, 8 * n, s));
    if (pos_zyx) BS_TRY(copy_out(pos_zyx, p->p->node_pos.p, 12 * n, s));
    if (sizes) BS_TRY(copy_out(sizes, p->p->node_size.p, 4 * n, s));
    return BS_OK;
}

int bs_stage1_block_counts(const bs_plan *p, int64_t *counts) {
    BS_ARG(p && counts, "null argument");
    for (size_t i = 0; i < p->p->blocks.size(); i++) counts[i] = p->p->blocks[i].owned ? p->p->block_count[i] : 0;
    return BS_OK;
}

int bs_stage1_set_block_counts(bs_plan *p, const int64_t *counts) {
    BS_ARG(p && counts, "null argument");
    Plan &P = *p->p;
    for (size_t i = 0; i < P.blocks.size(); i++) {
        if (P.blocks[i].owned && P.block_count[i] != 0 && P.block_count[i] != counts[i]) {
            set_error("bs_stage1_set_block_counts: counts of owned blocks differ from this rank's stage-1 result");
            return BS_ERR_STATE;
        }
        P.block_count[i] = counts[i];
    }
    P.block_nbase[0] = 0;
    for (size_t b = 0; b < P.blocks.size(); b++) P.block_nbase[b + 1] = P.block_nbase[b] + P.block_count[b];
    P.node_first = P.owned.empty() ? 0 : P.block_nbase[P.owned[0]];
    P.counts_global = true;
    return BS_OK;
}

int bs_plan_node_ids(bs_plan *p, uint64_t *ids_out, int64_t *n_out, void *stream) {
    BS_ARG(p, "null plan");
    Plan &P = *p->p;
    if (P.owned.size() != P.blocks.size() && !P.counts_global) {
        set_error("bs_plan_node_ids: multi-rank plan needs bs_stage1_set_block_counts first");
        return BS_ERR_STATE;
    }
    long long n = 0;
    BS_TRY(plan_node_ids(P, ids_out, &n, (cudaStream_t)stream));
    if (n_out) *n_out = n;
    return BS_OK;
}

int bs_stage2_agglomerate(bs_plan *p, const void *affs, const uint64_t *frags, void *stream) {
    BS_ARG(p && affs && frags, "bs_stage2_agglomerate: null argument");
    Plan &P = *p->p;
    if (P.owned.size() != P.blocks.size() && !P.counts_global) {
        set_error("bs_stage2_agglomerate: multi-rank plan needs bs_stage1_set_block_counts first");
        return BS_ERR_STATE;
    }
    return stage2_run(P, affs, frags, (cudaStream_t)stream, nullptr);
}

int bs_waterz_segment(bs_plan *p, const void *affs, const uint64_t *frags, const float *thresholds, int n_thresholds,
                      uint64_t *const *segs_out, uint32_t *counters_out, void *stream) {
    return bs_waterz_segment_quantile(p, affs, frags, thresholds, n_thresholds, 0, 0, segs_out, counters_out, stream);
}

int bs_waterz_segment_quantile(bs_plan *p, const void *affs, const uint64_t *frags, const float *thresholds, int n_thresholds,
                               int quantile, int init_with_max, uint64_t *const *segs_out, uint32_t *counters_out, void *stream) {
    BS_ARG(p && affs && frags && thresholds && segs_out && n_thresholds >= 1, "bs_waterz_segment: null argument");
    BS_ARG(quantile >= 0 && quantile < 100, "bs_waterz_segment_quantile: quantile must lie in 0..99");
    for (int t = 1; t < n_thresholds; t++)
        BS_ARG(thresholds[t] >= thresholds[t - 1], "bs_waterz_segment: thresholds must be ascending (waterz sorts them)");
    Plan &P = *p->p;
    BS_ARG(P.blocks.size() == 1 && P.owned.size() == 1, "bs_waterz_segment: needs a single-block plan (block = roi, context 0)");
    PqRequest rq;
    rq.thresholds = thresholds;
    rq.T = n_thresholds;
    rq.segs = segs_out;
    rq.quantile = quantile;
    rq.initmax = init_with_max ? 1 : 0;
    for (int i = 0; i < 4; i++) rq.counters[i] = 0;
    int rc = stage2_run(P, affs, frags, (cudaStream_t)stream, &rq);
    if (rc == BS_OK && counters_out)
        for (int i = 0; i < 4; i++) counters_out[i] = rq.counters[i];
    return rc;
}

int bs_stage2_num_edges(const bs_plan *p, int64_t *n) {
    BS_ARG(p && n, "null argument");
    *n = p->p->n_edges;
    return BS_OK;
}

int bs_stage2_get_edges(const bs_plan *p, uint64_t *u, uint64_t *v, float *score, void *stream) {
    BS_ARG(p, "null plan");
    cudaStream_t s = (cudaStream_t)stream;
    size_t n = (size_t)p->p->n_edges;
    if (u) BS_TRY(copy_out(u, p->p->edge_u.p, 8 * n, s));
    if (v) BS_TRY(copy_out(v, p->p->edge_v.p, 8 * n, s));
    if (score) BS_TRY(copy_out(score, p->p->edge_score.p, 4 * n, s));
    return BS_OK;
}

int bs_connected_components(const uint64_t *nodes, int64_t n, const uint64_t *edges_u, const uint64_t *edges_v,
                            const float *scores, int64_t m, float threshold, uint64_t *components_out, void *stream) {
    BS_ARG(n == 0 || (nodes && components_out), "bs_connected_components: null argument");
    BS_ARG(m == 0 || (edges_u && edges_v), "bs_connected_components: null edges");
    return connected_components(nodes, n, edges_u, edges_v, scores, m, threshold, components_out, (cudaStream_t)stream);
}

int bs_stage3_components(bs_plan *p, const uint64_t *nodes, int64_t n, const uint64_t *edges_u, const uint64_t *edges_v,
                         const float *scores, int64_t m, const float *thresholds, int n_thresholds, uint64_t *const *components_out,
                         void *stream) {
    BS_ARG(p && thresholds && components_out && (n == 0 || nodes), "bs_stage3_components: null argument");
    BS_ARG(m == 0 || (edges_u && edges_v && scores), "bs_stage3_components: null edges");
    return components_multi(*p->p, nodes, n, edges_u, edges_v, scores, m, thresholds, n_thresholds, components_out,
                            (cudaStream_t)stream);
}

int bs_relabel(const uint64_t *frags, int64_t n_vox, const uint64_t *lut_keys, const uint64_t *lut_vals, int64_t n_lut,
               uint64_t *seg_out, void *stream) {
    BS_ARG(n_vox == 0 || (frags && seg_out), "bs_relabel: null argument");
    BS_ARG(n_lut == 0 || (lut_keys && lut_vals), "bs_relabel: null lut");
    return relabel(frags, n_vox, lut_keys, lut_vals, n_lut, seg_out, (cudaStream_t)stream);
}

int bs_stage3_relabel(bs_plan *p, const uint64_t *frags, int64_t n_vox, const uint64_t *const *components, int n_thresholds,
                      uint64_t *const *segs_out, void *stream) {
    BS_ARG(p && (n_vox == 0 || (frags && components && segs_out)), "bs_stage3_relabel: null argument");
    return relabel_dense(*p->p, frags, n_vox, components, n_thresholds, segs_out, (cudaStream_t)stream);
}

int bs_stage1_from_labels(bs_plan *p, const void *affs, const uint8_t *mask, const uint32_t *labels, int64_t n_labels, uint64_t *frags_out,
                          void *stream) {
    BS_ARG(p && affs && labels && frags_out, "bs_stage1_from_labels: null argument");
    BS_ARG(n_labels >= 0 && n_labels < (1LL << 31), "bs_stage1_from_labels: label count out of range");
    init_mempool();
    return stage1_run(*p->p, affs, mask, frags_out, (cudaStream_t)stream, labels, (size_t)n_labels);
}

int bs_stage2_agglomerate_until(bs_plan *p, const void *affs, const uint64_t *frags, float threshold, void *stream) {
    BS_ARG(p && affs && frags, "bs_stage2_agglomerate_until: null argument");
    BS_ARG(threshold >= 0.0f && threshold <= 1.0f, "bs_stage2_agglomerate_until: threshold must lie in [0, 1]");
    init_mempool();
    p->p->agg_threshold = threshold;
    const int rc = stage2_run(*p->p, affs, frags, (cudaStream_t)stream, nullptr);
    p->p->agg_threshold = 1.0f;
    return rc;
}

int bs_stage3_dense_fragments(bs_plan *p, const uint64_t *frags, int64_t n_vox, uint32_t *dense_out, void *stream) {
    BS_ARG(p && (n_vox == 0 || (frags && dense_out)), "bs_stage3_dense_fragments: null argument");
    return dense_fragments(*p->p, frags, n_vox, dense_out, (cudaStream_t)stream);
}

#if defined(__x86_64__)
#include <emmintrin.h>
static inline void host_stream_store(uint64_t *p, uint64_t v) { _mm_stream_si64((long long *)p, (long long)v); }
static inline void host_stream_fence() { _mm_sfence(); }
#else
static inline void host_stream_store(uint64_t *p, uint64_t v) { *p = v; }
static inline void host_stream_fence() {}
#endif

// host side of the compact result form: dense ids + node-id table + LUT rows -> the uint64 arrays the reference writes
int bs_expand_compact(const uint32_t *dense, int64_t n_vox, const uint64_t *node_ids, int64_t n_nodes, const uint64_t *const *luts,
                      int n_thresholds, uint64_t *frags_out, uint64_t *const *segs_out, int n_threads) {
    BS_ARG(n_vox == 0 || (dense && node_ids), "bs_expand_compact: null argument");
    BS_ARG(n_thresholds >= 0 && n_thresholds <= 8, "bs_expand_compact: 0..8 thresholds");
    for (int t = 0; t < n_thresholds; t++) BS_ARG(luts && luts[t] && segs_out && segs_out[t], "bs_expand_compact: null LUT / output");
    const int nt = std::max(1, std::min(n_threads, 256));
    std::vector<std::thread> pool;
    std::atomic<int> bad(0);
    const int64_t chunk = (n_vox + nt - 1) / nt;
    for (int k = 0; k < nt; k++) {
        const int64_t a = (int64_t)k * chunk, b = std::min<int64_t>(n_vox, a + chunk);
        if (a >= b) break;
        pool.emplace_back([=, &bad]() {
            // the outputs are written once and not read back here: streaming stores skip the read-for-ownership of every
            // output line (the decoder is bound by host memory traffic)
            for (int64_t i = a; i < b; i++) {
                const uint32_t d = dense[i];
                if ((int64_t)d > n_nodes) {
                    bad.store(1);
                    continue;
                }
                if (frags_out) host_stream_store(&frags_out[i], d ? node_ids[d - 1] : 0);
                for (int t = 0; t < n_thresholds; t++) host_stream_store(&segs_out[t][i], d ? luts[t][d - 1] : 0);
            }
            host_stream_fence();
        });
    }
    for (auto &th : pool) th.join();
    BS_ARG(bad.load() == 0, "bs_expand_compact: a dense id exceeds the node table");
    return BS_OK;
}

int bs_watershed_from_affinities(const void *affs, int aff_dtype, int Z, int Y, int X, int fragments_in_xy,
                                 int min_seed_distance, uint64_t *frags_out, uint64_t *seeds_out, int64_t *n_out,
                                 void *stream) {
    BS_ARG(affs && frags_out, "bs_watershed_from_affinities: null argument");
    BS_ARG(seeds_out == nullptr, "bs_watershed_from_affinities: return_seeds is not implemented");
    bs_ws_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.vol_shape[0] = cfg.roi_shape[0] = cfg.block_size[0] = Z;
    cfg.vol_shape[1] = cfg.roi_shape[1] = cfg.block_size[1] = Y;
    cfg.vol_shape[2] = cfg.roi_shape[2] = cfg.block_size[2] = X;
    cfg.aff_dtype = aff_dtype;
    cfg.n_channels = 3;
    cfg.fragments_in_xy = fragments_in_xy;
    cfg.min_seed_distance = min_seed_distance;
    cfg.remove_debris = 0;
    cfg.filter_fragments = 0.0;
    cfg.queue_bins = 256;
    cfg.keep_cheaper = 1;
    cfg.crop_relabel = 1;
    cfg.block_begin = cfg.block_end = -1;
    Plan *p = nullptr;
    BS_TRY(plan_build(cfg, &p));
    int rc = stage1_run(*p, affs, nullptr, frags_out, (cudaStream_t)stream);
    if (rc == BS_OK && n_out) *n_out = p->n_nodes;
    cudaStreamSynchronize((cudaStream_t)stream);
    delete p;
    return rc;
}

int bs_cc_affs(const void *affs, int aff_dtype, const uint8_t *mask, int Z, int Y, int X, float threshold, int remove_debris,
               uint64_t *frags_out, uint64_t *seg_out, int64_t *n_out, void *stream) {
    BS_ARG(affs && frags_out, "bs_cc_affs: null argument");
    BS_ARG(aff_dtype == BS_DTYPE_U8 || aff_dtype == BS_DTYPE_F32, "bs_cc_affs: aff_dtype must be u8 or f32");
    init_mempool();
    return cc_affs(affs, aff_dtype, mask, Z, Y, X, threshold, remove_debris, frags_out, seg_out, n_out, (cudaStream_t)stream);
}

int bs_mws_agglom(const void *affs, int aff_dtype, const uint8_t *mask, int n_channels, int Z, int Y, int X, const int32_t *offsets,
                  const int32_t *strides, const double *bias, double noise_eps, unsigned long long noise_seed, int remove_debris,
                  uint64_t *frags_out, uint64_t *seg_out, int64_t *counters_out, void *stream) {
    BS_ARG(affs && offsets && frags_out, "bs_mws_agglom: null argument");
    BS_ARG(aff_dtype == BS_DTYPE_U8 || aff_dtype == BS_DTYPE_F32, "bs_mws_agglom: aff_dtype must be u8 or f32");
    init_mempool();
    return bs::mws_agglom(affs, aff_dtype, mask, n_channels, Z, Y, X, offsets, strides, bias, noise_eps, noise_seed, 1, remove_debris,
                          frags_out, seg_out, counters_out, (cudaStream_t)stream);
}

int bs_mws_agglom_blocks(const void *affs, int aff_dtype, const uint8_t *mask, int n_channels, int n_blocks, int Z, int Y, int X,
                         const int32_t *offsets, const int32_t *strides, const double *bias, double noise_eps, const uint64_t *block_seeds,
                         uint32_t *labels_out, int64_t *counters_out, void *stream) {
    BS_ARG(affs && offsets && labels_out, "bs_mws_agglom_blocks: null argument");
    BS_ARG(aff_dtype == BS_DTYPE_U8 || aff_dtype == BS_DTYPE_F32, "bs_mws_agglom_blocks: aff_dtype must be u8 or f32");
    BS_ARG(noise_eps == 0.0 || block_seeds, "bs_mws_agglom_blocks: noise needs one seed per block");
    init_mempool();
    return bs::mws_agglom_blocks(affs, aff_dtype, mask, n_channels, n_blocks, Z, Y, X, offsets, strides, bias, noise_eps,
                                 (const unsigned long long *)block_seeds, 1, labels_out, counters_out, (cudaStream_t)stream);
}

int bs_aff_agglom(bs_plan *p, const void *affs, const uint64_t *frags, int n_channels, const int32_t *offsets, void *stream) {
    BS_ARG(p && affs && frags && offsets, "bs_aff_agglom: null argument");
    Plan &P = *p->p;
    if (P.owned.size() != P.blocks.size() && !P.counts_global) {
        set_error("bs_aff_agglom: multi-rank plan needs bs_stage1_set_block_counts first");
        return BS_ERR_STATE;
    }
    BS_ARG(n_channels == P.cfg.n_channels, "bs_aff_agglom: n_channels differs from the plan's");
    init_mempool();
    return aff_agglom(P, affs, frags, n_channels, offsets, (cudaStream_t)stream);
}

int bs_graph_mws(const uint64_t *nodes, int64_t n, const uint64_t *edges_u, const uint64_t *edges_v, const float *scores, int64_t m,
                 double weight, double bias, uint64_t *clusters_out, int64_t *counters_out, void *stream) {
    BS_ARG(n == 0 || (nodes && clusters_out), "bs_graph_mws: null argument");
    BS_ARG(m == 0 || (edges_u && edges_v && scores), "bs_graph_mws: null edges");
    init_mempool();
    return bs::graph_mws(nodes, n, edges_u, edges_v, scores, m, weight, bias, clusters_out, counters_out, (cudaStream_t)stream);
}

int bs_label_stats(const uint64_t *seg, const int32_t *shape, int64_t capacity, uint64_t *ids_out, int64_t *sizes_out,
                   int32_t *zmin_out, int32_t *zmax_out, int64_t *n_out, void *stream) {
    init_mempool();
    return label_stats(seg, shape, capacity, ids_out, sizes_out, zmin_out, zmax_out, n_out, (cudaStream_t)stream);
}

int bs_aff_errors(const uint64_t *seg, const void *pred, int pred_dtype, int n_offsets, const int32_t *shape, const int32_t *neighborhood,
                  const uint8_t *mask, float floor_, float ceil_, float *seg_affs_out, float *error_map_out, uint8_t *error_mask_out,
                  void *stream) {
    init_mempool();
    return aff_errors(seg, pred, pred_dtype, n_offsets, shape, neighborhood, mask, floor_, ceil_, seg_affs_out, error_map_out,
                      error_mask_out, (cudaStream_t)stream);
}

int bs_shift_affinities(const void *affs, int aff_dtype, const uint8_t *mask, int Z, int Y, int X, const int32_t *radius,
                        const double *const *weights, const double *bias, float *out, void *stream) {
    BS_ARG(affs && out, "bs_shift_affinities: null argument");
    BS_ARG(aff_dtype == BS_DTYPE_U8 || aff_dtype == BS_DTYPE_F32, "bs_shift_affinities: aff_dtype must be u8 or f32");
    const int has_sigma = radius && weights && (radius[0] >= 0 || radius[1] >= 0 || radius[2] >= 0);
    if (has_sigma)
        for (int d = 0; d < 3; d++) BS_ARG(radius[d] < 0 || (radius[d] * 2 + 1 <= BS_SIGMA_MAXW && weights[d]), "bs_shift_affinities: bad gaussian kernel");
    init_mempool();
    return shift_affinities(affs, aff_dtype, mask, Z, Y, X, has_sigma, radius, weights, bias != nullptr, bias, out, (cudaStream_t)stream);
}

int bs_synth_affs(void *out, int aff_dtype, const int32_t *shape, const int32_t *offset, const int32_t *vol_shape,
                  uint64_t seed, void *stream) {
    BS_ARG(out && shape && offset, "bs_synth_affs: null argument");
    (void)vol_shape;
    return synth_affs(out, aff_dtype, shape, offset, seed, (cudaStream_t)stream);
}

int bs_debug_fetch(const bs_plan *p, const char *name, void *dst_host, int64_t *n_out) {
    BS_ARG(p && name, "null argument");
    auto it = p->p->dbg.find(name);
    if (it == p->p->dbg.end()) {
        set_error(std::string("bs_debug_fetch: no such array: ") + name);
        return BS_ERR_STATE;
    }
    auto meta = p->p->dbg_meta[name];
    if (n_out) *n_out = meta.second;
    if (dst_host) {
        BS_CUDA(cudaDeviceSynchronize());
        BS_CUDA(cudaMemcpy(dst_host, it->second->p, (size_t)meta.first * meta.second, cudaMemcpyDeviceToHost));
    }
    return BS_OK;
}

/* test hooks for the device primitives (prims.cu) */
int bs_dbg_scan_u32(const uint32_t *in, uint32_t *out, int64_t n, uint32_t *total_dev, void *stream) {
    return scan_exclusive_u32(in, out, (size_t)n, total_dev, (cudaStream_t)stream);
}
int bs_dbg_scan_u8(const uint8_t *in, uint32_t *out, int64_t n, uint32_t *total_dev, void *stream) {
    return scan_exclusive_u8(in, out, (size_t)n, total_dev, (cudaStream_t)stream);
}
int bs_dbg_sort_pairs(uint64_t *keys, uint32_t *vals, uint64_t *keys_tmp, uint32_t *vals_tmp, int64_t n, int bit_lo,
                      int bit_hi, void *stream) {
    return radix_sort_pairs(keys, vals, keys_tmp, vals_tmp, (size_t)n, bit_lo, bit_hi, (cudaStream_t)stream);
}

int bs_release_scratch(void) {
    int dev = 0;
    BS_CUDA(cudaGetDevice(&dev));
    cudaMemPool_t pool;
    BS_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    BS_CUDA(cudaDeviceSynchronize());
    g_arena.destroy();
    BS_CUDA(cudaMemPoolTrimTo(pool, 0));
    return BS_OK;
}

int bs_set_profiling(int on) {
    g_prof.on = on != 0;
    return BS_OK;
}

int bs_get_profile(char *names_out, int names_cap, float *ms_out, int cap, int *n_out) {
    int n = 0;
    std::string names;
    for (auto &r : g_prof.result) {
        if (n >= cap) break;
        if (ms_out) ms_out[n] = r.second;
        names += r.first;
        names += ";";
        n++;
    }
    if (names_out && names_cap > 0) {
        strncpy(names_out, names.c_str(), names_cap - 1);
        names_out[names_cap - 1] = 0;
    }
    if (n_out) *n_out = n;
    return BS_OK;
}

}  // extern "C"
