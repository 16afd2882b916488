// tiled.cuh — large grids (W a multiple of 32, any H): overlapped tiles + temporal blocking.
//
// A warp owns a 256 x 256-cell tile in registers (the same lane layout and generation code as
// the warp-resident batched kernel: lane L holds tile rows 8L..8L+7, 8 words each) and advances
// it T <= TV generations without touching memory.  The tile is treated as a small torus, which
// is wrong only within T cells of its border, so a halo of TV rows above/below and one 32-cell
// word left/right is discarded: each tile WRITES the interior (256 - 2*TV) rows x 192 columns
// and tiles overlap by the halo (224 columns when T <= 16: then only half of each edge word is
// stale and the exact halves are written with 16-bit stores).  One HBM/L2 round trip therefore covers T generations
// (algorithmic bytes stay 0.25 B per cell-generation; DRAM bytes drop by ~T).
//
// Two vertical modes:
//   torus  - single-GPU grid: tile rows are taken modulo H (toroidal wrap, carle/env.py:98-104)
//   band   - one GPU's row band of a giant grid: the local buffer holds
//            [TV halo rows | band rows | TV halo rows]; nothing wraps vertically, and the rows
//            that form the neighbours' halos are ALSO stored straight into the neighbouring
//            GPUs' buffers through peer-mapped pointers (NVLink), so the halo exchange is part
//            of the compute kernel instead of a separate copy.
// Horizontally the grid always wraps (word index modulo the row length).
#pragma once
#include "kernels.cuh"

namespace carle {

struct TiledParams {
    StepParams s;            // in/out, n, h (= local rows), w, wpr, window, flags, counters, k = T
    int tv;                  // vertical halo in rows, multiple of 8, >= T
    int out_row0, out_rows;  // rows [out_row0, out_row0 + out_rows) of the buffer are produced
    int vwrap;               // 1: torus (rows modulo h); 0: band (clamp, never needed)
    int act_row_shift;       // grid row of local row r is r + act_row_shift (band mode)
    int grid_h;              // rows of the whole torus (== s.h unless this is a band)
    int tiles_y, tiles_x;
    int xstride;             // grid words between tile origins: 6 (one discarded word per side)
                             // or 7 (T <= 16: only 16 columns per side are stale, so the two
                             // half words at the tile edges are written with 16-bit stores)
    uint32_t* peer_up;       // band mode: neighbour buffers (same layout) or nullptr
    uint32_t* peer_dn;
};

template <class Rule>
__global__ void __launch_bounds__(128, 2)
step_tiled_kernel(const __grid_constant__ TiledParams tp) {
    constexpr int WPR = 8;
    const StepParams& p = tp.s;
    const int lane = threadIdx.x & 31;
    const long long warps_per_block = blockDim.x >> 5;
    const long long warp0 = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * warps_per_block;
    const Rule rule(p);
    const int up_lane = (lane + 31) & 31, dn_lane = (lane + 1) & 31;
    const int interior_rows = 256 - 2 * tp.tv;
    const long long tiles_per_inst = (long long)tp.tiles_y * tp.tiles_x;
    const long long total_tiles = p.n * tiles_per_inst;
    const long long inst_words = (long long)p.h * p.wpr;

    for (long long tile = warp0; tile < total_tiles; tile += nwarps) {
        const long long inst = tile / tiles_per_inst;
        const int rem = (int)(tile - inst * tiles_per_inst);
        const int ty = rem / tp.tiles_x, tx = rem - ty * tp.tiles_x;
        const int tile_row0 = tp.out_row0 + ty * interior_rows - tp.tv;   // buffer row of tile row 0
        const int tile_word0 = tx * tp.xstride - 1;                       // grid word of tile word 0
        const uint32_t* src = p.in + inst * inst_words;

        // grid word index of each tile word (horizontal torus)
        int gw[WPR];
#pragma unroll
        for (int w = 0; w < WPR; ++w) {
            int g = (tile_word0 + w) % p.wpr;
            gw[w] = g < 0 ? g + p.wpr : g;
        }
        uint32_t x[WPR][WPR];
#pragma unroll
        for (int r = 0; r < WPR; ++r) {
            int row = tile_row0 + lane * WPR + r;
            if (tp.vwrap) { row %= p.h; if (row < 0) row += p.h; }
            else row = min(max(row, 0), p.h - 1);
            const uint32_t* rp = src + (long long)row * p.wpr;
#pragma unroll
            for (int w = 0; w < WPR; ++w) x[r][w] = rp[gw[w]];
        }

        for (int g = 0; g < p.k; ++g) {
            if (p.act) {
                // action XOR (carle/env.py:179-182) on whatever part of the window the tile holds
                const uint32_t* act_inst = p.act + (long long)g * p.act_step_stride +
                                           inst * p.act_inst_stride;
#pragma unroll
                for (int r = 0; r < WPR; ++r) {
                    int row = tile_row0 + lane * WPR + r;
                    if (tp.vwrap) { row %= p.h; if (row < 0) row += p.h; }
                    int grow = row + tp.act_row_shift;          // row of the whole torus
                    if (grow < 0) grow += tp.grid_h;
                    if (grow >= tp.grid_h) grow -= tp.grid_h;
                    const int ar = grow - p.row0;
                    if (ar >= 0 && ar < p.aw) {
#pragma unroll
                        for (int w = 0; w < WPR; ++w) {
                            const int j = gw[w] - p.aw0;
                            if (j >= 0 && j < p.awpr) x[r][w] ^= act_inst[(long long)ar * p.awpr + j];
                        }
                    }
                }
            }
            const bool reset = p.flags && p.flags[2 * g] == 0;
            if (reset) {
#pragma unroll
                for (int i = 0; i < WPR * WPR; ++i) (&x[0][0])[i] = 0u;
            } else {
                generation<WPR>(x, rule, up_lane, dn_lane);
            }
        }

        // ---- write the interior: lanes [tv/8, 32 - tv/8), tile words 1..6 ----
        const int lane0 = tp.tv >> 3;
        if (lane >= lane0 && lane < 32 - lane0) {
            uint32_t* dst = p.out + inst * inst_words;
#pragma unroll
            for (int r = 0; r < WPR; ++r) {
                const int orow = tile_row0 + lane * WPR + r;          // buffer row (no wrap needed:
                if (orow >= tp.out_row0 + tp.out_rows) continue;      //  out rows are inside it)
                uint32_t* rp = dst + (long long)orow * p.wpr;
                // band mode: the first / last tv produced rows are the neighbours' halos
                uint32_t* up = nullptr;
                uint32_t* dn = nullptr;
                if (tp.peer_up && orow < tp.out_row0 + tp.tv)
                    up = tp.peer_up + inst * inst_words +
                         (long long)(orow + tp.out_rows) * p.wpr;
                if (tp.peer_dn && orow >= tp.out_row0 + tp.out_rows - tp.tv)
                    dn = tp.peer_dn + inst * inst_words +
                         (long long)(orow - tp.out_rows) * p.wpr;
#pragma unroll
                for (int w = 1; w < WPR - 1; ++w) {
                    if (tx * tp.xstride + (w - 1) >= p.wpr) continue; // partial last tile column
                    rp[gw[w]] = x[r][w];
                    if (up) up[gw[w]] = x[r][w];
                    if (dn) dn[gw[w]] = x[r][w];
                }
                if (tp.xstride == 7) {
                    // after T <= 16 generations bits 16..31 of tile word 0 and bits 0..15 of
                    // tile word 7 are still exact: they complete the neighbouring tiles' words
                    const uint16_t hi0 = (uint16_t)(x[r][0] >> 16);
                    const uint16_t lo7 = (uint16_t)(x[r][WPR - 1] & 0xFFFFu);
                    reinterpret_cast<uint16_t*>(rp + gw[0])[1] = hi0;
                    if (up) reinterpret_cast<uint16_t*>(up + gw[0])[1] = hi0;
                    if (dn) reinterpret_cast<uint16_t*>(dn + gw[0])[1] = hi0;
                    if (tx * 7 + 6 < p.wpr) {
                        reinterpret_cast<uint16_t*>(rp + gw[WPR - 1])[0] = lo7;
                        if (up) reinterpret_cast<uint16_t*>(up + gw[WPR - 1])[0] = lo7;
                        if (dn) reinterpret_cast<uint16_t*>(dn + gw[WPR - 1])[0] = lo7;
                    }
                }
            }
        }
    }
    retire_block(p);
}

}  // namespace carle
