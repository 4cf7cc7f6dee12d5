// Block / tile geometry of one blockwise `bs segment --ws` run (host + device PODs) and the
// plan object behind the C ABI.  Internal header.
#pragma once
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/bsnative.h"
#include "common.cuh"

namespace bs {

// One watershed problem: a 2-D slice (xy mode, D = 1, ndim = 2) or a whole 3-D read ROI
// (ndim = 3) of a daisy block.  Per-pixel scratch arrays are indexed base + (z*H + y)*W + x.
struct Tile {
    int gz, gy, gx;      // dataset coordinates of tile voxel (0,0,0); may lie outside the volume (zero fill)
    int D, H, W;         // tile extent
    int wz, wy, wx;      // write region offset inside the tile
    int wD, wH, wW;      // write region extent
    int block;           // index into the plan's block table
    int ndim;            // 2 or 3 (array rank the reference hands to scipy / skimage)
    long long base;      // offset of this tile in the per-pixel scratch arrays (batch local)
    long long wbase;     // offset of this tile's write region in batch write order
    FastDiv fW, fH, fwW, fwH;   // divisions by W, H, wW, wH
    void set_divs() {
        fW = make_fastdiv((uint32_t)W), fH = make_fastdiv((uint32_t)H);
        fwW = make_fastdiv((uint32_t)wW), fwH = make_fastdiv((uint32_t)wH);
    }
};

// 1-D gaussian kernels of the sigma shift as scipy builds them (device pointers), per spatial axis; radius < 0: skipped
struct GaussWeights {
    int radius[3];
    const double *w[3];
};

struct Blk {
    long long block_id;      // daisy block id (cantor number of the block index)
    int idx[3];              // block grid index
    int wo[3], ws[3];        // write ROI offset (dataset voxel coords) and shape
    int ro[3], rs[3];        // read ROI offset and shape (write grown by context; not clipped)
    int nb[27];              // plan block indices of the 3x3x3 neighbourhood (self at 13), -1 if none
    int owned;               // 1 if this rank processes the block
    int pad_;
};

struct Plan {
    bs_ws_config cfg;
    std::vector<Blk> blocks;             // ALL blocks of the task, ascending block_id
    std::vector<int> owned;              // indices of owned blocks (ascending)
    long long nvox_block;                // prod(block_size)
    // ---- stage 1 results
    std::vector<long long> block_count;  // fragments per block (all blocks once counts are known)
    std::vector<long long> block_nbase;  // exclusive prefix of block_count (dense node numbering)
    bool counts_global = false;
    DevBuf fidx;                         // u32 roi_shape: dense node index + 1 (0 = background) -- owned blocks only
    DevBuf node_id, node_pos, node_size; // owned nodes, ascending id
    long long n_nodes = 0;
    long long node_first = 0;            // dense index of the first owned node
    // ---- stage 2 results
    DevBuf edge_u, edge_v, edge_score;   // owned edges
    long long n_edges = 0;
    float agg_threshold = 1.0f;          // waterz mergeUntil threshold of the next stage-2 call (1.0: the blockwise path; epsilon_agglomerate: eps)
    // ---- debug scratch kept from the last run (name -> buffer)
    std::map<std::string, DevBuf *> dbg;
    std::map<std::string, std::pair<int, long long>> dbg_meta;  // name -> (element bytes, count)
    ~Plan();
};

int plan_build(const bs_ws_config &cfg, Plan **out);
// fragment id -> dense node number map of the plan's current fragment counts (table in `buf`, uploaded on `s`)
int plan_idmap(const Plan &P, DevBuf &buf, IdMap *idm, cudaStream_t s);

}  // namespace bs
