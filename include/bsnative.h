/*
 * libbsnative — C ABI of the B200-native `bs segment --ws` hot path.
 *
 * This is the drop-in boundary: plain C types, device pointers owned by the caller
 * (torch tensors on the Python side), a cudaStream_t passed as void*.  Every entry
 * point returns 0 on success or a negative BS_ERR_* code; bs_last_error() gives the
 * thread-local message.  There is NO CPU fallback anywhere behind this ABI.
 * Scratch arena, launch counter and stage profiler are per host thread (thread_local; the arena also per device): several
 * host threads may drive different plans concurrently, one plan is driven by one thread at a time.  Calls are ordered on
 * the stream they are given and synchronise it where a count has to reach the host.
 *
 * Each entry point names the reference interface it replaces (paths relative to
 * /root/reference/bootstrapper, upstream ucsdmanorlab/bootstrapper v0.3.2).
 */
#ifndef BSNATIVE_H
#define BSNATIVE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BS_OK 0
#define BS_ERR_CUDA (-1)
#define BS_ERR_ARG (-2)
#define BS_ERR_OVERFLOW (-3)
#define BS_ERR_STATE (-4)

#define BS_SIGMA_MAXW 129   /* gaussian kernels up to radius 64 (sigma <= 16) */
#define BS_DTYPE_U8 0
#define BS_DTYPE_F32 1

/* Resolved `ws_params` + blockwise geometry (segment.py:11-23 DEFAULTS["ws"],
 * post/watershed.py:28-98).  All coordinates are voxels, (z, y, x). */
typedef struct bs_ws_config {
    int32_t vol_shape[3];      /* spatial shape of the affinity array (C, Z, Y, X)          */
    int32_t roi_offset[3];     /* task ROI inside the array (post/watershed.py:69-73)       */
    int32_t roi_shape[3];
    int32_t block_size[3];     /* block_shape (post/watershed.py:75-78)                      */
    int32_t context[3];        /* context     (post/watershed.py:79-83)                      */
    int32_t aff_dtype;         /* BS_DTYPE_U8 (normalised /255) or BS_DTYPE_F32              */
    int32_t n_channels;        /* C >= 3; only [:3] is read (watershed_frags.py:117)         */
    int32_t fragments_in_xy;   /* ws_params.fragments_in_xy                                  */
    int32_t min_seed_distance; /* ws_params.min_seed_distance                                */
    int32_t remove_debris;     /* ws_params.remove_debris (0 = off)                          */
    int32_t queue_bins;        /* waterz discretize_queue: 256 blockwise (waterz_agglom.py:136) */
    int32_t keep_cheaper;      /* oracle switch U6c (1 = default)                            */
    int32_t crop_relabel;      /* 1 = blockwise path: crop + skimage.measure.label + global ids
                                  (watershed_frags.py:216-224); 0 = raw watershed ids           */
    int32_t block_begin;       /* this rank owns blocks [block_begin, block_end) of the      */
    int32_t block_end;         /*   z-major block grid, in units of z-layers of blocks; -1/-1 = all */
    int32_t win_z0;            /* multi-GPU slab window: when win_z > 0 the affinity array is  */
    int32_t win_z;             /*   (C, win_z, Y, X) and the fragment array (win_z, Y, X), both holding
                                    the global planes [win_z0, win_z0 + win_z); roi must span all of z */
    double filter_fragments;   /* ws_params.filter_fragments (0 = off)                       */
    int64_t max_batch_voxels;  /* scratch bound for stage 1 (0 = default)                    */
    /* optional shifts of the affinities the watershed sees (watershed_frags.py:118-139):       */
    int32_t has_bias;          /* ws_params.bias given                                        */
    int32_t has_seed_eps;      /* ws_params.seed_eps given                                    */
    double bias[3];            /* per channel (a scalar bias is replicated by the caller)    */
    double seed_eps;           /* shift -= seed_eps * EDT(seeds == 0)                        */
    /* ws_params.sigma: shift += gaussian_filter(affs, (0, *sigma)) - affs.  The caller passes scipy's kernel:   */
    int32_t has_sigma;
    int32_t sigma_radius[3];   /* per axis (z, y, x): int(4 * sigma + 0.5); -1 = sigma of that axis <= 1e-15 */
    double sigma_w[3][BS_SIGMA_MAXW]; /* 2 * radius + 1 normalised weights per axis                          */
    /* daisy block ids: block.block_id[1] = cantor_number(write_roi.offset / write_roi.shape) with ABSOLUTE
     * offsets.  block_index_offset = the task ROI's absolute offset in voxels (dataset offset / voxel_size + roi_offset);
     * the grid index that is numbered is floor((block_index_offset + i * block_size) / block_size) per axis.           */
    int32_t block_index_offset[3];
    /* ws_params.noise_eps (watershed_frags.py:119-120: shift += np.random.randn(*affs.shape) * noise_eps, an UNSEEDED draw per
     * block upstream): a seeded counter-based generator stands in -- unit-variance sum of four 16-bit uniforms from a SplitMix64
     * hash of (noise_seed, block id, channel, raveled read-ROI voxel) -- mirrored bit for bit by the oracle */
    int32_t has_noise;
    double noise_eps;
    uint64_t noise_seed;
} bs_ws_config;

typedef struct bs_plan bs_plan;

const char *bs_last_error(void);
/* number of kernels launched by this library in this process (bench.py gpu_launches) */
unsigned long long bs_launch_count(void);
int bs_version(void);
/* sizeof(bs_ws_config) as this library was compiled: a binding checks its struct layout against it */
unsigned long long bs_config_size(void);

/* ---- plan: geometry of one `bs segment --ws -b` run ------------------------------- */
/* replaces: volara BlockwiseTask geometry as used by WatershedFrags / WaterzAgglom
 * (post/blockwise/watershed_frags.py:75-96, waterz_agglom.py:77-96) + daisy block
 * enumeration (blockwise.py:31-62). */
int bs_plan_create(const bs_ws_config *cfg, bs_plan **out);
void bs_plan_destroy(bs_plan *p);
int bs_plan_num_blocks(const bs_plan *p, int64_t *n_total, int64_t *n_owned);
/* per block (all blocks of the task, ascending daisy block id): block_id (cantor number),
 * write offset[3], write shape[3]; arrays of n_total entries (host pointers). */
int bs_plan_block_info(const bs_plan *p, int64_t *block_id, int32_t *write_offset, int32_t *write_shape);

/* restrict the blocks this plan processes to the given plan indices (ascending block-id order, as
 * returned by bs_plan_block_info): one daisy block = one process_block(block) call
 * (watershed_frags.py:248-258), or one rank's slab. */
int bs_plan_set_owned(bs_plan *p, const int32_t *indices, int64_t n);

/* ---- stage 1: fragments ------------------------------------------------------------
 * replaces: WatershedFrags.process_block for every owned block
 * (post/blockwise/watershed_frags.py:196-246), which itself calls
 * watershed_from_affinities (post/ws.py:38-112), filter_avg_fragments (:148-156),
 * skimage remove_small_objects (:188-192), skimage.measure.label (:222).
 *   affs      device, (C, Z, Y, X) of cfg.aff_dtype, C-contiguous
 *   mask      device, (Z, Y, X) uint8 or NULL (mask_dataset, watershed_frags.py:207-213)
 *   frags_out device, roi_shape uint64; owned blocks' write ROIs are written, the rest untouched
 */
int bs_stage1_fragments(bs_plan *p, const void *affs, const uint8_t *mask, uint64_t *frags_out, void *stream);
/* RAG nodes written by stage 1 (watershed_frags.py:230-246): id, position (voxels, write
 * offset + truncated centre of mass), size.  Host or device destination pointers. */
int bs_stage1_num_nodes(const bs_plan *p, int64_t *n);
int bs_stage1_get_nodes(const bs_plan *p, uint64_t *ids, int32_t *pos_zyx, uint32_t *sizes, void *stream);
/* per-block fragment counts of the owned blocks (n_total entries, 0 for blocks not owned) */
int bs_stage1_block_counts(const bs_plan *p, int64_t *counts);
/* multi-GPU: install the fragment counts of ALL blocks (after an all-gather) so that
 * stage 2 can number halo fragments of neighbouring ranks. */
int bs_stage1_set_block_counts(bs_plan *p, const int64_t *counts);

/* ids of ALL fragments of the task in ascending order (the key row of the fragment->segment LUT,
 * post/watershed.py:156-161,187) derived from the per-block counts; ids_out device (NULL: count only). */
int bs_plan_node_ids(bs_plan *p, uint64_t *ids_out, int64_t *n_out, void *stream);

/* ---- stage 2: RAG extraction + waterz agglomeration + merge-tree scores ------------
 * replaces: WaterzAgglom.process_block for every owned block
 * (post/blockwise/waterz_agglom.py:106-170): funlib.segment relabel (:116),
 * waterz.agglomerate(thresholds=[0,1], discretize_queue=256, merge history + region
 * graph) (:131-151), MergeTree replay + LCA query (:153-168, post/merge_tree.py),
 * write_graph ownership (:170).
 *   frags     device, roi_shape uint64 (the stage-1 output incl. neighbour ranks' halo)
 */
int bs_stage2_agglomerate(bs_plan *p, const void *affs, const uint64_t *frags, void *stream);
int bs_stage2_num_edges(const bs_plan *p, int64_t *n);
/* edges persisted by the owned blocks: u < v (fragment ids), merge_score (NaN = NULL) */
int bs_stage2_get_edges(const bs_plan *p, uint64_t *u, uint64_t *v, float *score, void *stream);

/* ---- epsilon_agglomerate (watershed_frags.py:158-176, 182-183) -------------------------------------------------------
 * The reference merges a block's watershed fragments with waterz (mean affinity, BinQueue<256>) up to a low threshold before it
 * filters, crops and relabels them.  Two entry points let a host driver do the same with the kernels above
 * (bootstrapper_b200/post/pipeline.py:segment_blockwise):
 *   bs_stage2_agglomerate_until  bs_stage2_agglomerate with waterz's mergeUntil(threshold) instead of the blockwise 1.0: edges
 *                                of merged pairs carry their merge score, all others NaN -- the connected components of the
 *                                scored edges are the merged fragments;
 *   bs_stage1_from_labels        the back half of bs_stage1_fragments (filter_avg_fragments, remove_small_objects, crop,
 *                                skimage.measure.label, ids, nodes) on GIVEN fragments: `labels` holds one (rz, ry, rx) uint32
 *                                volume per owned block (its read ROI, ascending block order, packed back to back), values 0
 *                                or 1..n_labels, unique over the call.  The blocks are treated as 3-D arrays (the given
 *                                fragments may span z slices). */
int bs_stage2_agglomerate_until(bs_plan *p, const void *affs, const uint64_t *frags, float threshold, void *stream);
int bs_stage1_from_labels(bs_plan *p, const void *affs, const uint8_t *mask, const uint32_t *labels, int64_t n_labels, uint64_t *frags_out,
                          void *stream);

/* ---- single-shot path: waterz with the default (non-discretised) queue ------------------
 * replaces: waterz.agglomerate(affs, thresholds, fragments=..., scoring_function=OneMinus<MeanAffinity>)
 * as simple_watershed drives it (post/watershed.py:333-340): region graph of the whole ROI, priority
 * queue ordered by (score, edge id), one segmentation per (ascending) threshold.
 *   plan       single block (block_size = roi_shape, context 0) whose stage 1 produced `frags`
 *   thresholds host, ascending;  segs_out host array of n_thresholds device pointers (roi_shape uint64)
 *   counters_out host[4] or NULL: pops, stale re-scores, deleted pops, merges
 */
int bs_waterz_segment(bs_plan *p, const void *affs, const uint64_t *frags, const float *thresholds, int n_thresholds,
                      uint64_t *const *segs_out, uint32_t *counters_out, void *stream);
/* the same call with one of the histogram scoring functions of post/watershed.py:232-244,
 * scoring_function=OneMinus<HistogramQuantileAffinity<RegionGraphType, Q, ScoreValue, 256, init_with_max>>
 * ("hist_quant_Q" / "hist_quant_Q_initmax"): quantile = Q in 1..99 (the reference offers 10, 25, 50, 75, 90);
 * quantile = 0 is bs_waterz_segment.  1 KB of device scratch per region-graph edge. */
int bs_waterz_segment_quantile(bs_plan *p, const void *affs, const uint64_t *frags, const float *thresholds, int n_thresholds,
                               int quantile, int init_with_max, uint64_t *const *segs_out, uint32_t *counters_out, void *stream);

/* ---- stage 3: global thresholded connected components + relabel --------------------
 * replaces: funlib.segment.graphs.impl.connected_components (post/watershed.py:182),
 * volara LUT (post/watershed.py:187-188) and volara Relabel (post/watershed.py:192-202).
 *   nodes (n) uint64 ascending, edges_u/v (m), scores (m, NaN skipped), all device
 *   components_out (n) uint64 device: component id = smallest node id of the component
 */
int bs_connected_components(const uint64_t *nodes, int64_t n, const uint64_t *edges_u, const uint64_t *edges_v,
                            const float *scores, int64_t m, float threshold, uint64_t *components_out,
                            void *stream);
/* the same for up to 8 (ascending) thresholds in one pass over the edges; `nodes` = bs_plan_node_ids of the plan (node
 * numbers come from the id arithmetic, no search); thresholds host, components_out host array of device pointers. */
int bs_stage3_components(bs_plan *p, const uint64_t *nodes, int64_t n, const uint64_t *edges_u, const uint64_t *edges_v,
                         const float *scores, int64_t m, const float *thresholds, int n_thresholds, uint64_t *const *components_out,
                         void *stream);
/* seg[i] = lut_vals[k] if frags[i] == lut_keys[k] else frags[i]; lut_keys ascending */
int bs_relabel(const uint64_t *frags, int64_t n_vox, const uint64_t *lut_keys, const uint64_t *lut_vals,
               int64_t n_lut, uint64_t *seg_out, void *stream);

/* Relabel for up to 8 thresholds in one pass over the fragments (volara Relabel, post/watershed.py:192-202):
 * components[t] is the LUT value row for threshold t, indexed by node in ascending-id order (all nodes of the
 * task, i.e. the plan must know every block's fragment count); ids unknown to the plan are copied through.
 * `components` / `segs_out` are HOST arrays of device pointers. */
int bs_stage3_relabel(bs_plan *p, const uint64_t *frags, int64_t n_vox, const uint64_t *const *components, int n_thresholds,
                      uint64_t *const *segs_out, void *stream);

/* ---- compact results (not in the reference: a transport form for callers on the far side of PCIe) ----------------------
 * Fragments and all T segmentations of a volume are functions of ONE 32-bit plane: dense[i] = 1 + the rank of frags[i] in the
 * ascending node list (0 = background), plus the node-id table (bs_plan_node_ids) and the T LUT rows of bs_stage3_components:
 *     frags[i] = node_ids[dense[i] - 1],   seg_t[i] = components_t[dense[i] - 1].
 * 4 bytes per voxel cross the bus instead of 8 (T + 1).  bs_expand_compact rebuilds the uint64 arrays on the HOST (plain
 * host pointers, n_threads worker threads); it is a decoder of this library's own format, not a CPU implementation of
 * the path. */
int bs_stage3_dense_fragments(bs_plan *p, const uint64_t *frags, int64_t n_vox, uint32_t *dense_out, void *stream);
int bs_expand_compact(const uint32_t *dense, int64_t n_vox, const uint64_t *node_ids, int64_t n_nodes, const uint64_t *const *luts,
                      int n_thresholds, uint64_t *frags_out, uint64_t *const *segs_out, int n_threads);

/* ---- `bs segment --cc` -----------------------------------------------------------------
 * replaces: cc_affs + compute_connected_component_segmentation (post/connected_components.py:15-127,
 * post/cc.py:7-74): hard = affs[:3] > threshold (float32 compare; uint8 input is /255 first), components of
 * the "+e_d" affinity graph, ids in raster order of each component's first voxel; remove_debris > 0 additionally
 * writes the remove_small_objects result to seg_out.
 *   affs (>=3, Z, Y, X) u8 / f32 (channels 0..2 read); mask (Z,Y,X) u8 or NULL; frags_out, seg_out (Z,Y,X) uint64
 *   (seg_out may be NULL); n_out (host) = number of components. */
int bs_cc_affs(const void *affs, int aff_dtype, const uint8_t *mask, int Z, int Y, int X, float threshold, int remove_debris,
               uint64_t *frags_out, uint64_t *seg_out, int64_t *n_out, void *stream);

/* ---- `bs segment --mws`: mutex watershed fragments --------------------------------------------
 * replaces: mwatershed_from_affinities + mwatershed.agglom (post/mws.py:12-59) as simple_mutex drives it
 * (post/watershed_mutex.py:177-291): weights = affs (uint8 / 255 in float64, or float32 widened; * (mask > 0)) + shift with
 * shift = noise * noise_eps + bias[c]; every (offset c, voxel p) with p + offset_c inside the volume and p on the stride
 * lattice of c is an edge; mutex watershed over the edges by descending |w| (ties: ascending (channel, raveled voxel),
 * declared deviation D4 -- the upstream sort is unstable); remove_debris > 0 additionally writes the
 * remove_small_objects result to seg_out.
 *   affs (C, Z, Y, X) u8 / f32; mask (Z,Y,X) u8 or NULL; offsets, strides (or NULL) host int32 [C*3]; bias host double [C];
 *   noise_eps != 0: a seeded counter-based generator stands in for the reference's unseeded np.random.randn (same
 *   generator in the oracle); sigma shifts and randomized_strides are not implemented.
 *   frags_out, seg_out (may be NULL) device (Z,Y,X) uint64: label = 1 + smallest raveled voxel index of the cluster;
 *   counters_out host[5] or NULL: edges, merges, mutex edges, attractive edges blocked by a mutex, rounds. */
int bs_mws_agglom(const void *affs, int aff_dtype, const uint8_t *mask, int n_channels, int Z, int Y, int X, const int32_t *offsets,
                  const int32_t *strides, const double *bias, double noise_eps, unsigned long long noise_seed, int remove_debris,
                  uint64_t *frags_out, uint64_t *seg_out, int64_t *counters_out, void *stream);

/* ---- `bs segment --mws -b`: the blockwise mutex-watershed pipeline --------------------------------
 * replaces the arithmetic of the volara tasks post/watershed_mutex.py:8-174 runs (ExtractFrags :123-141, AffAgglom :144-154,
 * GraphMWS :156-162; volara is third-party and not in the reference tree: restated, parity unpinned).
 *
 * bs_mws_agglom_blocks -- ExtractFrags' fragmenter for many blocks in one call: affs (C, n_blocks * Z, Y, X) holds the read
 *   ROIs (Z, Y, X) of n_blocks blocks stacked along z (zero fill / mask already applied by the caller as the task does); one
 *   independent mutex watershed per block exactly as bs_mws_agglom computes it (edges never leave a block; noise, if any,
 *   from block_seeds[block] -- host uint64 [n_blocks]).  labels_out (n_blocks * Z, Y, X) uint32: clusters numbered 1..n in the
 *   order of their first voxel in the stacked volume; counters_out host[6]: edges, merges, mutex edges, blocked, rounds, n.
 *   The rest of the task (filter, crop, relabel, id bump, nodes) is bs_stage1_from_labels.
 * bs_aff_agglom -- AffAgglom(scores={"zyx_aff": neighborhood}) for the plan's owned blocks: per block, fragments and
 *   affinities of its read ROI (zero fill outside); for every offset c and voxel p with p + offset_c inside the read ROI a pair
 *   of different non-zero fragments adds affs[c][p] to the edge (min id, max id); edge attribute = mean over all contributions
 *   (exact integer sums, one rounding); an edge is written by the block that owns its smaller endpoint.  Results: the plan's
 *   edge arrays (bs_stage2_num_edges / bs_stage2_get_edges), sorted by (u, v).
 * bs_graph_mws -- GraphMWS(weights={"zyx_aff": (weight, bias)}): mutex watershed on the fragment graph; w = weight * score +
 *   bias in float64 (NaN scores skipped), edges by descending |w| (equal |w|: input order -- pass edges sorted by (u, v)),
 *   w > 0 attractive else repulsive.  nodes ascending; clusters_out[i] = smallest node id of node i's cluster (the LUT row).
 *   counters_out host[5] or NULL: edges, merges, mutex edges, blocked, rounds. */
int bs_mws_agglom_blocks(const void *affs, int aff_dtype, const uint8_t *mask, int n_channels, int n_blocks, int Z, int Y, int X,
                         const int32_t *offsets, const int32_t *strides, const double *bias, double noise_eps, const uint64_t *block_seeds,
                         uint32_t *labels_out, int64_t *counters_out, void *stream);
int bs_aff_agglom(bs_plan *p, const void *affs, const uint64_t *frags, int n_channels, const int32_t *offsets, void *stream);
int bs_graph_mws(const uint64_t *nodes, int64_t n, const uint64_t *edges_u, const uint64_t *edges_v, const float *scores, int64_t m,
                 double weight, double bias, uint64_t *clusters_out, int64_t *counters_out, void *stream);

/* ---- per-label statistics (`bs refine`) --------------------------------------
 * replaces: the tile scans of refine.py -- `_global_sizes` (:98-108, fastremap.unique + bincount) and the z-extent loop
 * of `z_filter` (:236-255): voxel count, first and last z plane of every non-zero id of a label volume, ids ascending
 * (the order np.unique gives the reference).  The filters' decisions stay host arithmetic on these tables; masking and
 * remapping (`fastremap.mask` / `fastremap.remap(preserve_missing_labels=True)`, :111-116, :265-270) are bs_relabel.
 *   seg (Z,Y,X) uint64; *_out device arrays of `capacity` entries; n_out (host) = number of ids.
 *   BS_ERR_OVERFLOW when the volume holds more than `capacity` distinct ids (retry with a larger capacity). */
int bs_label_stats(const uint64_t *seg, const int32_t *shape, int64_t capacity, uint64_t *ids_out, int64_t *sizes_out,
                   int32_t *zmin_out, int32_t *zmax_out, int64_t *n_out, void *stream);

/* ---- affinity self-consistency error (`bs evaluate`) ------------------------
 * replaces: the compute of AddAffErrors.process (gp/add_aff_errors.py:128-183) on one array: seg -> affinities on
 * `neighborhood` (gunpowder seg_to_affgraph: same id, both > 0, 0 where the neighbour is outside), float32
 * error = sum_c (seg_aff - pred)^2 in channel order, * mask, / max, error_mask = floor < error < ceil.
 *   seg (Z,Y,X) uint64; pred (C,Z,Y,X) float32 (BS_DTYPE_F32) or uint8 (BS_DTYPE_U8, normalised * 1/255 as gp.Normalize
 *   does, eval/compute_errors.py:146); shape[3], neighborhood[C*3] host; mask (Z,Y,X) u8 or NULL;
 *   seg_affs_out (C,Z,Y,X) float32 or NULL; error_map_out (Z,Y,X) float32; error_mask_out (Z,Y,X) uint8. */
int bs_aff_errors(const uint64_t *seg, const void *pred, int pred_dtype, int n_offsets, const int32_t *shape, const int32_t *neighborhood,
                  const uint8_t *mask, float floor_, float ceil_, float *seg_affs_out, float *error_map_out, uint8_t *error_mask_out,
                  void *stream);

/* ---- shifts of the single-shot paths ----------------------------------------------------
 * replaces the numpy / scipy block of simple_watershed and cc_affs (post/watershed.py:262-303,
 * connected_components.py:52-77): out = affs_data + shift in float32, with affs_data = affs[:3].astype(float32)
 * (/ 255 for uint8) * (mask > 0) and shift = (gaussian_filter(affs_data, (0, *sigma)) - affs_data) + bias.
 *   radius[3] / weights[3] (host): scipy's kernel per axis as in bs_ws_config (NULL / -1 = no sigma); bias[3] or NULL;
 *   out device (3, Z, Y, X) float32. */
int bs_shift_affinities(const void *affs, int aff_dtype, const uint8_t *mask, int Z, int Y, int X, const int32_t *radius,
                        const double *const *weights, const double *bias, float *out, void *stream);

/* ---- post/ws.py plug point ----------------------------------------------------------
 * replaces: watershed_from_affinities(affs, max_affinity_value, fragments_in_xy,
 * return_seeds, min_seed_distance) (post/ws.py:38-112) on one in-memory array.
 *   affs (3, Z, Y, X) u8 (max_affinity_value 255) or f32 (1.0); frags_out (Z,Y,X) uint64;
 *   seeds_out optional (NULL).  n_out (host) = number of fragment ids used (max id).
 */
int bs_watershed_from_affinities(const void *affs, int aff_dtype, int Z, int Y, int X, int fragments_in_xy,
                                 int min_seed_distance, uint64_t *frags_out, uint64_t *seeds_out,
                                 int64_t *n_out, void *stream);

/* ---- test / bench harness (not part of the reference's path) -------------------------
 * seeded block-addressable synthetic affinities (same arithmetic as
 * bootstrapper_b200/synth.py), written for region [offset, offset+shape) of a volume. */
int bs_synth_affs(void *out, int aff_dtype, const int32_t *shape, const int32_t *offset, const int32_t *vol_shape,
                  uint64_t seed, void *stream);
/* debug: copy a named scratch array of the last stage-1 batch / stage-2 run to the host
 * (returns element count via n_out if dst is NULL). */
int bs_debug_fetch(const bs_plan *p, const char *name, void *dst_host, int64_t *n_out);
/* per-stage device times (ms) of the last run, measured with CUDA events on the caller's
 * stream when enabled via bs_set_profiling(1). names/values up to cap entries. */
int bs_set_debug(int on);
/* flood kernel: 0 = automatic (v2 for 2-D tiles up to 2^17 pixels, else the CTA-per-tile kernel v3), 1 = one-warp
 * global-memory flood (v1) everywhere, 2 = v2 with the tile bitmap in shared memory, 3 = v2 with the tile bitmap in
 * global memory (every tile resident at once), 4 = 3 + level tails in shared memory, 6 = the FAITHFUL flood: skimage's binary heap
 * replayed literally, one thread per tile (seed ties as the reference resolves them: no deviation D1; ~10^2 slower) */
int bs_set_flood_version(int v);
/* stage-1 front end (mask ... priority levels): 0 = automatic (the fused on-chip kernels for 2-D tiles of unshifted affinities
 * that fit one CTA's shared memory, else the unfused chain), 1 = unfused chain, 2 = fused with vector loads only,
 * 3 = fused, the TMA mask kernel required where the affinity rows are 16-byte aligned */
int bs_set_front_version(int v);
/* agglomeration kernel: 0 = automatic (parallel merges; shared memory when a block's graph fits, else a global slab),
 * 3 = parallel merges on global slabs for every block (queue bins, parents and stamps in shared memory), 4 = parallel merges
 * with every array in the global slab */
int bs_set_agglom_version(int v);
/* return the library's cached scratch memory (stream-ordered pool) to the driver */
int bs_release_scratch(void);
/* test hooks for the device primitives (exclusive scan, stable LSD radix sort) */
int bs_dbg_scan_u32(const uint32_t *in, uint32_t *out, int64_t n, uint32_t *total_dev, void *stream);
int bs_dbg_scan_u8(const uint8_t *in, uint32_t *out, int64_t n, uint32_t *total_dev, void *stream);
int bs_dbg_sort_pairs(uint64_t *keys, uint32_t *vals, uint64_t *keys_tmp, uint32_t *vals_tmp, int64_t n, int bit_lo,
                      int bit_hi, void *stream);
int bs_set_profiling(int on);
int bs_get_profile(char *names_out, int names_cap, float *ms_out, int cap, int *n_out);

#ifdef __cplusplus
}
#endif
#endif /* BSNATIVE_H */
