// fastvim_b200 -- token tiles of the full-resolution kernels (gate forward / backward, conv backward).
//
// A tile is TT <= 8 consecutive SEQUENCE positions of one image with all channels.  On plain
// (outer, pool, 1) geometries a tile is an equal split of one pooled group (tiles_per_group > 0),
// so every token of a tile shares one pooled position; channel layouts (inner > 1) use plain runs
// of TT positions (tiles_per_group == 0).  One warp fills the table; the hot loops then contain no
// integer division and no token -> row arithmetic.
#pragma once
#include "common.cuh"

namespace fv {

int sm_count();

// per-tile table in shared memory
template <int TT>
struct TileTab {
    long long yoff[TT];  // element offset of each token's row in y (row * ldy), image offset included
    int rows[TT + 6];    // memory token row of sequence positions t0-3 .. t0+TT+2 (-1: outside)
    int jt[TT];          // pooled position of each token
    int b, np, valid, pad;
};

template <int TT>
__device__ __forceinline__ void fill_tiletab(const Geom& g, int64_t tile, int64_t ntiles, int tiles_per_img,
                                             int tiles_per_group, int tile_len, int64_t ldy, int64_t ybs,
                                             TileTab<TT>* tab) {
    // executed by warp 0 only
    const int lane = threadIdx.x;
    if (tile >= ntiles) {
        if (lane == 0) tab->valid = 0;
        return;
    }
    const int b = (int)(tile / tiles_per_img), rem = (int)(tile - (int64_t)b * tiles_per_img);
    int t0, np;
    if (tiles_per_group > 0) {
        const int j = rem / tiles_per_group, q = rem - j * tiles_per_group;
        t0 = j * g.pool + q * tile_len;
        np = min(tile_len, g.pool - q * tile_len);
    } else {
        t0 = rem * TT;
        np = min(TT, g.L - t0);
    }
    if (lane < TT + 6) {
        const int t = t0 - 3 + lane;
        const int row = (lane < np + 6 && t >= 0 && t < g.L) ? (int)seq_to_row(g, t) : -1;
        tab->rows[lane] = row;
        if (lane >= 3 && lane < TT + 3) tab->yoff[lane - 3] = (int64_t)b * ybs + (int64_t)(row < 0 ? 0 : row) * ldy;
    }
    if (lane < TT) tab->jt[lane] = lane < np ? seq_to_pooled(g, t0 + lane) : 0;
    if (lane == 0) {
        tab->b = b;
        tab->np = np;
        tab->valid = 1;
    }
}


}  // namespace fv
