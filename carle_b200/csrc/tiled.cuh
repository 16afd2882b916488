// tiled.cuh — large grids (W a multiple of 32, any H): overlapped tiles + temporal blocking.
//
// A warp owns a (32*R) x 256-cell tile in registers (the same lane layout and generation code as
// the warp-resident batched kernel: lane L holds tile rows R*L..R*L+R-1, 8 words each; R = 8:
// 256 rows, 255 registers, 8 warps per SM; R = 4: 128 rows, 16 warps per SM) and advances
// it T <= TV generations without touching memory.  The tile is treated as a small torus, which
// is wrong only within T cells of its border, so a halo of TV rows above/below and one 32-cell
// word left/right is discarded: each tile WRITES the interior (32*R - 2*TV) rows x 192 columns
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
    int tv;                  // vertical halo in rows, multiple of 8 (rows per lane divide it), >= T
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
    // band mode, neighbour-only synchronisation of consecutive temporal blocks WITHOUT a host or
    // NCCL barrier (nullptr: the caller synchronises).  sync_local = this rank's int32[4]:
    // [0] blocks the UPPER neighbour has finished, [1] blocks the LOWER neighbour has finished
    // (both written remotely, by those neighbours), [2] blocks this rank has finished.  A launch
    // first waits until both neighbours have finished as many blocks as this rank (their edge
    // rows have landed in this rank's halos, and they no longer read the halos this launch is
    // about to overwrite), and its last CTA publishes the new count in the neighbours' words.
    int* sync_local;
    int* sync_up_word;       // &upper neighbour's sync_local[1] (this rank is its lower neighbour)
    int* sync_dn_word;       // &lower neighbour's sync_local[0]
};

__device__ __forceinline__ int ld_acquire_sys(const int* ptr) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int* ptr, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" :: "l"(ptr), "r"(v) : "memory");
}

// resident CTAs (4 warps each) asked of ptxas per rows-per-lane
constexpr int tiled_min_ctas(int r) { return r >= 8 ? 2 : 4; }
// dynamic shared memory of a 4-warp CTA: two padded transposition slabs per warp
constexpr int tiled_smem_bytes(int r) { return 4 * 2 * 32 * (r * 8 + 1) * 4; }

template <class Rule, int R>
__global__ void __launch_bounds__(128, tiled_min_ctas(R))
step_tiled_kernel(const __grid_constant__ TiledParams tp) {
    constexpr int WPR = 8;                      // words per tile row
    static_assert(R == 4 || R == 8, "rows per lane");
    const StepParams& p = tp.s;
    const int lane = threadIdx.x & 31;
    const long long warps_per_block = blockDim.x >> 5;
    // (warp index through a shuffle: the compiler then knows the tile loop is warp-uniform and
    //  drops the WARPSYNC / ENDCOLLECTIVE pair it otherwise wraps around every neighbour shuffle)
    const long long warp0 = (long long)blockIdx.x * warps_per_block +
                            __shfl_sync(0xFFFFFFFFu, (int)(threadIdx.x >> 5), 0);
    const long long nwarps = (long long)gridDim.x * warps_per_block;
    const Rule rule(p);
    const int up_lane = (lane + 31) & 31, dn_lane = (lane + 1) & 31;
    const int interior_rows = 32 * R - 2 * tp.tv;
    const long long tiles_per_inst = (long long)tp.tiles_y * tp.tiles_x;
    const long long total_tiles = p.n * tiles_per_inst;
    const long long inst_words = (long long)p.h * p.wpr;

    // Memory side.  A tile row is 8 words = 32 bytes, and lane L owns rows R*L..R*L+R-1: loading or
    // storing "my rows" directly makes every warp instruction touch 32 different sectors for 4
    // bytes each (the LSU then spends 32 cycles per instruction: ~1 ms per launch on a 65536^2
    // grid, as much as eight generations of arithmetic).  Instead the warp moves FOUR WHOLE ROWS
    // per instruction (lane = row-in-group * 8 + word: 4 to 8 sectors) and transposes through a
    // padded shared-memory slab [owner lane][R*8 + 1] (conflict free on both sides):
    //   load   cp.async (4 bytes per lane) global -> slab A, issued one trip ahead, so the next
    //          tile arrives while this one is being advanced;
    //   store  registers -> slab B -> coalesced stores (32-bit interior words, 16-bit halves of
    //          the two edge words when T <= 16).
    constexpr int STRIDE = R * WPR + 1;
    constexpr int GROUPS = 32 * R / 4;                        // four tile rows per warp instruction
    extern __shared__ uint32_t tile_smem[];
    const int wib = threadIdx.x >> 5;
    uint32_t* slab_a = tile_smem + (size_t)(2 * wib) * 32 * STRIDE;
    uint32_t* slab_b = slab_a + 32 * STRIDE;
    const int rg = lane >> 3, wl = lane & 7;                  // this lane's row-in-group and word
    if (tp.sync_local) {
        // neighbour-only barrier between temporal blocks (see TiledParams::sync_local)
        if (threadIdx.x == 0) {
            const int done = *reinterpret_cast<volatile int*>(tp.sync_local + 2);
            while (ld_acquire_sys(tp.sync_local + 0) < done) {}
            while (ld_acquire_sys(tp.sync_local + 1) < done) {}
        }
        __syncthreads();
    }
    auto prefetch = [&](long long tile) {
        const long long inst = tile / tiles_per_inst;
        const int rem = (int)(tile - inst * tiles_per_inst);
        const int ty = rem / tp.tiles_x, tx = rem - ty * tp.tiles_x;
        const int tile_row0 = tp.out_row0 + ty * interior_rows - tp.tv;
        int gwl = (tx * tp.xstride - 1 + wl) % p.wpr;         // grid word of tile word wl (torus)
        if (gwl < 0) gwl += p.wpr;
        const uint32_t* src = p.in + inst * inst_words + gwl;
        int row = tile_row0 + rg;                             // buffer row of this lane's first row
        if (tp.vwrap) { row %= p.h; if (row < 0) row += p.h; }
        // pointers advance by four rows per copy; the slab offset alternates between +4 rows inside
        // a lane's slab and the step into the next lane's slab (as in the store loop below)
        const uint32_t* rp = src + (long long)row * p.wpr;
        const uint32_t* last = src + (long long)(p.h - 1) * p.wpr;    // band mode: clamp (never used)
        const long long step = 4LL * p.wpr, wrap_back = (long long)p.h * p.wpr;
        uint32_t sdst = tma::smem_u32(slab_a + rg * WPR + wl);
        constexpr int S0 = (R == 8) ? 4 * WPR : STRIDE;
        constexpr int S1 = (R == 8) ? STRIDE - 4 * WPR : STRIDE;
        if (row + 4 * (GROUPS - 1) < p.h) {
            // the tile's rows neither wrap around the torus nor leave the band buffer (all but the
            // tiles on the grid's last tile row): plain pointer stepping, 3 instructions per copy
            // instead of the 17 of the general loop (65536^2: 1090 -> 200 per tile)
#pragma unroll 8
            for (int i = 0; i < GROUPS; ++i) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(sdst), "l"(rp) : "memory");
                rp += step;
                sdst += 4u * ((i & 1) ? S1 : S0);
            }
        } else {
#pragma unroll 8
        for (int i = 0; i < GROUPS; ++i) {
            const uint32_t* q = (!tp.vwrap && row >= p.h) ? last : rp;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(sdst), "l"(q) : "memory");
            row += 4;
            rp += step;
            if (tp.vwrap && row >= p.h) { row -= p.h; rp -= wrap_back; }   // (tiles are never taller than the grid)
            sdst += 4u * ((i & 1) ? S1 : S0);
        }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (warp0 < total_tiles) prefetch(warp0);

    for (long long tile = warp0; tile < total_tiles; tile += nwarps) {
        const long long inst = tile / tiles_per_inst;
        const int rem = (int)(tile - inst * tiles_per_inst);
        const int ty = rem / tp.tiles_x, tx = rem - ty * tp.tiles_x;
        const int tile_row0 = tp.out_row0 + ty * interior_rows - tp.tv;   // buffer row of tile row 0
        const int tile_word0 = tx * tp.xstride - 1;                       // grid word of tile word 0

        uint32_t x[R][WPR];
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();                                   // every lane's copies have landed
#pragma unroll
        for (int i = 0; i < R * WPR; ++i) (&x[0][0])[i] = slab_a[lane * STRIDE + i];

        if (!p.act && !p.flags) {
            // free run (no actions, no reset flags): nothing but generations.  The next tile's copies go
            // out first (slab A was read above), and the loop is unrolled by two so that the register
            // renaming of one generation is undone by the next instead of by ~30 moves per generation
            __syncwarp();
            if (tile + nwarps < total_tiles) prefetch(tile + nwarps);
            int g = 0;
#pragma unroll 1
            for (; g + 2 <= p.k; g += 2) {
                generation_rw<R, WPR>(x, rule, up_lane, dn_lane);
                generation_rw<R, WPR>(x, rule, up_lane, dn_lane);
            }
            if (g < p.k) generation_rw<R, WPR>(x, rule, up_lane, dn_lane);
        } else
        for (int g = 0; g < p.k; ++g) {
            if (p.act) {
                // action XOR (carle/env.py:179-182) on whatever part of the window the tile holds
                const uint32_t* act_inst = p.act + (long long)g * p.act_step_stride +
                                           inst * p.act_inst_stride;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    int row = tile_row0 + lane * R + r;
                    if (tp.vwrap) { row %= p.h; if (row < 0) row += p.h; }
                    int grow = row + tp.act_row_shift;          // row of the whole torus
                    if (grow < 0) grow += tp.grid_h;
                    if (grow >= tp.grid_h) grow -= tp.grid_h;
                    const int ar = grow - p.row0;
                    if (ar >= 0 && ar < p.aw) {
#pragma unroll
                        for (int w = 0; w < WPR; ++w) {
                            int gw = (tile_word0 + w) % p.wpr;  // grid word of tile word w (torus)
                            if (gw < 0) gw += p.wpr;
                            const int j = gw - p.aw0;
                            if (j >= 0 && j < p.awpr) x[r][w] ^= act_inst[(long long)ar * p.awpr + j];
                        }
                    }
                }
            }
            const bool reset = p.flags && p.flags[2 * g] == 0 && !p.defer_reset;
            if (reset) {
#pragma unroll
                for (int i = 0; i < R * WPR; ++i) (&x[0][0])[i] = 0u;
            } else {
                generation_rw<R, WPR>(x, rule, up_lane, dn_lane);
            }
            // (behind the first generation: slab A has long been read by every lane)
            if (g == 0 && tile + nwarps < total_tiles) {
                __syncwarp();
                prefetch(tile + nwarps);
            }
        }

        // ---- write the interior: tile rows [tv, 32R - tv), tile words 1..6 (+ edge halves) ----
        __syncwarp();                                   // slab B: the previous tile's reads are done
#pragma unroll
        for (int i = 0; i < R * WPR; ++i) slab_b[lane * STRIDE + i] = (&x[0][0])[i];
        __syncwarp();
        {
            int gwl = (tile_word0 + wl) % p.wpr;
            if (gwl < 0) gwl += p.wpr;
            uint32_t* dst = p.out + inst * inst_words + gwl;
            // which part of tile word wl this lane stores: 2 = all 32 bits, 1 = bits 16..31 (tile
            // word 0), 3 = bits 0..15 (tile word 7), 0 = nothing.  After T <= 16 generations those
            // halves of the edge words are still exact: they complete the neighbouring tiles' words.
            int part = 0;
            if (wl >= 1 && wl <= WPR - 2) part = (tx * tp.xstride + (wl - 1) < p.wpr) ? 2 : 0;
            else if (tp.xstride == 7) part = (wl == 0) ? 1 : ((tx * 7 + 6 < p.wpr) ? 3 : 0);
            const int g_lo = tp.tv >> 2, g_hi = (32 * R - tp.tv) >> 2;       // interior row groups
            const int out_end = tp.out_row0 + tp.out_rows;
            // band mode: only the tiles that produce the band's first / last tv rows also store
            // into the neighbouring GPUs' buffers; they take the general loop below
            const bool edge_tile = (tp.peer_up && tile_row0 < tp.out_row0) ||
                                   (tp.peer_dn && tile_row0 + 32 * R - tp.tv > out_end - tp.tv);
            if (!edge_tile) {
                // fast path: pointers advance by four rows per trip; the slab offset alternates
                // between +4 rows inside a lane's slab and the step into the next lane's slab
                int i_end = (out_end - tile_row0 - rg + 3) >> 2;      // rows past the grid's / band's end
                i_end = part ? min(i_end, g_hi) : g_lo;
                const int trow0 = 4 * g_lo + rg;
                const uint32_t* sp = slab_b + (trow0 / R) * STRIDE + (trow0 % R) * WPR + wl;
                uint32_t* rp = dst + (long long)(tile_row0 + trow0) * p.wpr;
                const long long step = 4LL * p.wpr;
                constexpr int S0 = (R == 8) ? 4 * WPR : STRIDE;       // slab step of an even / odd group
                constexpr int S1 = (R == 8) ? STRIDE - 4 * WPR : STRIDE;
                if (part == 2 && i_end >= g_hi) {
                    // every interior row of the tile lies inside the produced rows (all tiles but the
                    // last tile row's): no per-row test
#pragma unroll 4
                    for (int i = g_lo; i < g_hi; i += 2) {            // (g_hi - g_lo is even)
                        *rp = sp[0];
                        rp[step] = sp[S0];
                        rp += 2 * step;
                        sp += S0 + S1;
                    }
                } else if (part == 2) {
#pragma unroll 2
                    for (int i = g_lo; i < g_hi; i += 2) {            // (g_hi - g_lo is even)
                        if (i < i_end) *rp = sp[0];
                        if (i + 1 < i_end) rp[step] = sp[S0];
                        rp += 2 * step;
                        sp += S0 + S1;
                    }
                } else {
                    uint16_t* hp = reinterpret_cast<uint16_t*>(rp) + ((part == 1) ? 1 : 0);
                    const int sh = (part == 1) ? 16 : 0;
#pragma unroll 2
                    for (int i = g_lo; i < g_hi; i += 2) {
                        if (i < i_end) *hp = (uint16_t)(sp[0] >> sh);
                        if (i + 1 < i_end) hp[2 * step] = (uint16_t)(sp[S0] >> sh);
                        hp += 4 * step;
                        sp += S0 + S1;
                    }
                }
            } else
            for (int i = (part ? g_lo : g_hi); i < g_hi; ++i) {
                const int trow = 4 * i + rg;
                const int orow = tile_row0 + trow;                    // buffer row (no wrap needed:
                if (orow >= tp.out_row0 + tp.out_rows) continue;      //  out rows are inside it)
                const uint32_t v = slab_b[(trow / R) * STRIDE + (trow % R) * WPR + wl];
                uint32_t* rp = dst + (long long)orow * p.wpr;
                // band mode: the first / last tv produced rows are the neighbours' halos
                uint32_t* up = nullptr;
                uint32_t* dn = nullptr;
                if (tp.peer_up && orow < tp.out_row0 + tp.tv)
                    up = tp.peer_up + inst * inst_words + gwl + (long long)(orow + tp.out_rows) * p.wpr;
                if (tp.peer_dn && orow >= tp.out_row0 + tp.out_rows - tp.tv)
                    dn = tp.peer_dn + inst * inst_words + gwl + (long long)(orow - tp.out_rows) * p.wpr;
                if (part == 2) {
                    *rp = v;
                    if (up) *up = v;
                    if (dn) *dn = v;
                } else {
                    const int half = (part == 1) ? 1 : 0;
                    const uint16_t hv = (part == 1) ? (uint16_t)(v >> 16) : (uint16_t)(v & 0xFFFFu);
                    reinterpret_cast<uint16_t*>(rp)[half] = hv;
                    if (up) reinterpret_cast<uint16_t*>(up)[half] = hv;
                    if (dn) reinterpret_cast<uint16_t*>(dn)[half] = hv;
                }
            }
        }
    }
    if (!tp.sync_local) {
        retire_block(p);
        return;
    }
    // band mode with neighbour flags: every CTA makes its stores (also the ones into the
    // neighbours' buffers, over NVLink) visible system-wide before it retires; the last CTA does
    // the bookkeeping and publishes "one more block finished" in both neighbours' words
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        if (atomicAdd(p.retire, 1u) == gridDim.x - 1) {
            __threadfence_system();
            finish_step(p);
            *p.retire = 0u;
            const int done = tp.sync_local[2] + 1;
            tp.sync_local[2] = done;
            st_release_sys(tp.sync_up_word, done);
            st_release_sys(tp.sync_dn_word, done);
        }
    }
}

}  // namespace carle
