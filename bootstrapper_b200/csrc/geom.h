// Block / tile geometry shared by the stage kernels (host + device PODs).
#pragma once
#include <stdint.h>

namespace bs {

// One watershed problem: a 2-D slice (xy mode, D = 1) or a 3-D block (D > 1) of a
// daisy block's read ROI.  Per-pixel scratch arrays are indexed base + (z*H + y)*W + x.
struct Tile {
    int gz, gy, gx;      // dataset coordinates of tile voxel (0,0,0) (may be negative / beyond: zero fill)
    int D, H, W;         // tile extent
    int wz, wy, wx;      // write region offset inside the tile
    int wD, wH, wW;      // write region extent
    int block;           // linear block index
    int pad_;
    long long base;      // offset of this tile in the per-pixel scratch arrays
    long long wbase;     // offset of this tile's write region in block-raster write order
};

struct Blk {
    long long block_id;      // daisy block id (cantor number)
    int wo[3], ws[3];        // write ROI (dataset voxel coords), shape
    int ro[3], rs[3];        // read ROI
    int nb[27];              // linear index of the 27 neighbouring blocks (incl. self at 13), -1 if none
    int tile_first, tile_count;
    long long wbase;         // first write-order index of the block
};

struct VolGeom {
    int Z, Y, X;             // dataset (affinity array) spatial shape
    int ro[3], rs[3];        // task ROI offset / shape inside the dataset (fragments array has shape rs)
    int bs[3];               // block size
    int ctx[3];              // context
    int nb[3];               // blocks per axis
};

}  // namespace bs
