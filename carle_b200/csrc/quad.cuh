// quad.cuh — one env step of 256 x 256 instances with FOUR warps per instance.
// (Superseded by the independent strips of strip.cuh -- 206 us vs 106 us per step at config 3 --
//  and kept as the CARLE_FUSED_IMPL=quad variant for A/B runs.)
//
// The one-warp-per-instance kernels hold a whole 256 x 256 universe in 64 registers per lane
// (255 in total): 8 warps per SM, issue-bound at a third of the HBM roofline.  Here a group of
// four warps shares one instance that the TMA engine has staged in shared memory (read-only old
// state, 8 KiB, double-buffered): warp q advances the 64-row band q (lane L holds rows 2L, 2L+1,
// eight words each = 16 state registers), so ~3x more warps are resident and every phase of one
// warp overlaps with other warps' memory waits.  Cross-band coupling is small and goes through
// shared memory: the 64 x 64 action (staged by the same bulk copy) is ballotted by all four
// warps (16 window rows each),
// the post-action row triples of each band's first / last row are exchanged for the vertical
// neighbours, and the four partial SpeedDetector sums are combined by warp 0.  Three named
// barriers (128 threads) per instance.
#pragma once
#include "kernels.cuh"

namespace carle {

template <typename T>
struct QuadGroupSmem {
    uint32_t state[2][2048];        // two TMA slots: packed 256 x 256 universe
    T act[2][4096];                 //                unpacked 64 x 64 action of the same instance
    uint32_t amask[64][2];          // ballotted action rows (window row r, 32-column chunk c)
    uint32_t edge[4][2][8][2];      // [warp][first/last row][word][lo/hi plane] row triples
    uint32_t sums[4][4];            // partial live, sh, sw, window-live per warp
    unsigned long long full[2];     // mbarriers of the two slots
};

__device__ __forceinline__ void group_sync(int group) {
    asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory");
}

#ifndef CARLE_QUAD_CTAS
#define CARLE_QUAD_CTAS 2
#endif

template <class Rule, typename T>
__global__ void __launch_bounds__(256, CARLE_QUAD_CTAS)
step_quad_kernel(const __grid_constant__ StepParams p) {
    constexpr int WPL = 8;                       // words per row
    extern __shared__ __align__(128) unsigned char quad_smem_raw[];
    __shared__ unsigned int s_done;
    __shared__ int s_flag[3];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int group = wib >> 2, q = wib & 3;     // q: band of the instance this warp advances
    QuadGroupSmem<T>& sm = reinterpret_cast<QuadGroupSmem<T>*>(quad_smem_raw)[group];
    const long long ngroups = (long long)gridDim.x * 2;
    const long long gid = (long long)blockIdx.x * 2 + group;

    if (threadIdx.x == 0) { s_done = 0u; s_flag[0] = 0; s_flag[1] = 0; s_flag[2] = 0; }
    if (q == 0 && lane == 0) {
        tma::mbar_init(reinterpret_cast<uint64_t*>(&sm.full[0]), 1);
        tma::mbar_init(reinterpret_cast<uint64_t*>(&sm.full[1]), 1);
        tma::fence_mbar_init();
    }
    __syncthreads();

    const Rule rule(p);
    const char* in_bytes = reinterpret_cast<const char*>(p.in);
    const T* act_base = static_cast<const T*>(p.raw);
    auto issue = [&](int s, long long inst) {    // one thread of the group
        uint64_t* bar = reinterpret_cast<uint64_t*>(&sm.full[s]);
        tma::mbar_expect_tx(bar, 8192 + 4096 * (uint32_t)sizeof(T));
        tma::bulk_g2s(sm.state[s], in_bytes + inst * 8192, 8192, bar);
        tma::bulk_g2s(sm.act[s], act_base + inst * p.raw_inst_stride, 4096 * (uint32_t)sizeof(T), bar);
    };
    if (gid < p.n && q == 0 && lane == 0) issue(0, gid);

    bool warp_not_one = false, warp_any = false, warp_nonbin = false;
    const int bit0 = p.col0 - 32 * p.aw0;        // 0 for the 256/64 geometry, kept general
    uint32_t phase0 = 0u, phase1 = 0u;           // mbarrier parities of the two slots
    int it = 0;
    for (long long inst = gid; inst < p.n; inst += ngroups, ++it) {
        const int s = it & 1;
        // (slot s^1 was released by the barrier that ended the previous trip)
        const long long next = inst + ngroups;
        if (next < p.n && q == 0 && lane == 0) issue(s ^ 1, next);

        tma::mbar_wait(reinterpret_cast<uint64_t*>(&sm.full[s]), s ? phase1 : phase0);
        if (s) phase1 ^= 1u; else phase0 ^= 1u;
        // ---- action: this warp ballots window rows [16q, 16q+16), two 32-column chunks ----
        const T* a = &sm.act[s][(16 * q) * 64 + lane];
        T v[16][2];
#pragma unroll
        for (int r = 0; r < 16; ++r) { v[r][0] = a[r * 64]; v[r][1] = a[r * 64 + 32]; }
        uint32_t differs = 0u, seen = 0u, mine = 0u;
        NonBinary nb;
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, v[r][c] != T(0));
                differs |= bits_of(v[r][c]) ^ OneBits<T>::value;
                nb.see(v[r][c]);
                seen |= m;
                if (lane == 2 * r + c) mine = m;
            }
        sm.amask[16 * q + (lane >> 1)][lane & 1] = mine;          // lane 2r+c holds row r, chunk c
        warp_not_one |= __any_sync(0xFFFFFFFFu, differs != 0u);
        warp_any |= (seen != 0u);
        warp_nonbin |= __any_sync(0xFFFFFFFFu, nb.any_lane());
        group_sync(group);                                        // (1) action masks visible

        // ---- this lane's two rows of band q ----
        uint32_t x[2][WPL];
        const int row_a = 64 * q + 2 * lane;                      // instance row of x[0]
        {
            const uint4* src = reinterpret_cast<const uint4*>(&sm.state[s][row_a * WPL]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint4 t = src[i];
                (&x[0][0])[4 * i + 0] = t.x; (&x[0][0])[4 * i + 1] = t.y;
                (&x[0][0])[4 * i + 2] = t.z; (&x[0][0])[4 * i + 3] = t.w;
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int ar = row_a + r - p.row0;                    // window row, if any
            if (ar >= 0 && ar < p.aw) {
                const uint32_t m0 = sm.amask[ar][0], m1 = sm.amask[ar][1];
                const uint32_t w0 = m0 << bit0;
                const uint32_t w1 = bit0 ? ((m1 << bit0) | (m0 >> (32 - bit0))) : m1;
                const uint32_t w2 = bit0 ? (m1 >> (32 - bit0)) : 0u;
#pragma unroll
                for (int w = 0; w < WPL; ++w) {
                    if (w == p.aw0) x[r][w] ^= w0;
                    if (w == p.aw0 + 1) x[r][w] ^= w1;
                    if (w == p.aw0 + 2) x[r][w] ^= w2;
                }
            }
        }
        // ---- row triples; the band's first / last row go to the neighbouring warps ----
        ca::Triple t0[WPL], t1[WPL], up[WPL], dn[WPL];
#pragma unroll
        for (int w = 0; w < WPL; ++w) {
            const int wl = (w + WPL - 1) % WPL, wr = (w + 1) % WPL;
            t0[w] = ca::row_triple(ca::west(x[0][wl], x[0][w]), x[0][w], ca::east(x[0][w], x[0][wr]));
            t1[w] = ca::row_triple(ca::west(x[1][wl], x[1][w]), x[1][w], ca::east(x[1][w], x[1][wr]));
        }
        if (lane == 0) {
#pragma unroll
            for (int w = 0; w < WPL; ++w) { sm.edge[q][0][w][0] = t0[w].lo; sm.edge[q][0][w][1] = t0[w].hi; }
        }
        if (lane == 31) {
#pragma unroll
            for (int w = 0; w < WPL; ++w) { sm.edge[q][1][w][0] = t1[w].lo; sm.edge[q][1][w][1] = t1[w].hi; }
        }
        group_sync(group);                                        // (2) edge triples visible
#pragma unroll
        for (int w = 0; w < WPL; ++w) {
            up[w].lo = __shfl_up_sync(0xFFFFFFFFu, t1[w].lo, 1);
            up[w].hi = __shfl_up_sync(0xFFFFFFFFu, t1[w].hi, 1);
            dn[w].lo = __shfl_down_sync(0xFFFFFFFFu, t0[w].lo, 1);
            dn[w].hi = __shfl_down_sync(0xFFFFFFFFu, t0[w].hi, 1);
            if (lane == 0) {                                      // last row of the band above
                up[w].lo = sm.edge[(q + 3) & 3][1][w][0];
                up[w].hi = sm.edge[(q + 3) & 3][1][w][1];
            }
            if (lane == 31) {                                     // first row of the band below
                dn[w].lo = sm.edge[(q + 1) & 3][0][w][0];
                dn[w].hi = sm.edge[(q + 1) & 3][0][w][1];
            }
        }
#pragma unroll
        for (int w = 0; w < WPL; ++w) {
            const uint32_t n0 = rule.from_triples(x[0][w], up[w], t0[w], t1[w]);
            const uint32_t n1 = rule.from_triples(x[1][w], t0[w], t1[w], dn[w]);
            x[0][w] = n0;
            x[1][w] = n1;
        }
        // ---- next state: 64 contiguous bytes per lane ----
        {
            uint4* dst = reinterpret_cast<uint4*>(p.out + inst * 2048 + row_a * WPL);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                dst[i] = make_uint4((&x[0][0])[4 * i + 0], (&x[0][0])[4 * i + 1],
                                    (&x[0][0])[4 * i + 2], (&x[0][0])[4 * i + 3]);
        }
        // ---- fused SpeedDetector sums (carle/mcl.py:773-779): partial per warp ----
        if (p.red) {
            uint32_t live = 0, sh = 0, sw = 0, wl = 0;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int row = row_a + r;
                const bool in_rows = (row >= p.row0) && (row < p.row0 + p.aw);
                uint32_t rowcnt = 0;
#pragma unroll
                for (int w = 0; w < WPL; ++w) {
                    const uint32_t val = x[r][w];
                    const uint32_t inside = in_rows ? (val & window_col_mask(p, w)) : 0u;
                    const uint32_t outside = val ^ inside;
                    const uint32_t c = ca::popc32(outside);
                    live += ca::popc32(val);
                    wl += ca::popc32(inside);
                    rowcnt += c;
                    sw += 32u * w * c + ca::bit_index_sum(outside);
                }
                sh += (uint32_t)row * rowcnt;
            }
            live = __reduce_add_sync(0xFFFFFFFFu, live);
            sh = __reduce_add_sync(0xFFFFFFFFu, sh);
            sw = __reduce_add_sync(0xFFFFFFFFu, sw);
            wl = __reduce_add_sync(0xFFFFFFFFu, wl);
            if (lane == 0) { sm.sums[q][0] = live; sm.sums[q][1] = sh; sm.sums[q][2] = sw; sm.sums[q][3] = wl; }
        }
        group_sync(group);                                        // (3) slot s and smem reusable
        if (p.red && q == 0 && lane < 4) {
            const unsigned long long tot = (unsigned long long)sm.sums[0][lane] + sm.sums[1][lane] +
                                           sm.sums[2][lane] + sm.sums[3][lane];
            p.red[inst * 4 + lane] = (long long)tot;
        }
    }
    // ---- retirement: warp -> block (shared memory) -> grid (global) ----
    if (retire_legacy<T>(p, &s_done, s_flag, lane, (int)(blockDim.x >> 5), warp_not_one, warp_any,
                         warp_nonbin) == 2)
        clear_after_reset(p, lane);
}

}  // namespace carle
