// strip.cuh — one env step with an instance split into independent STRIPS of 32*R rows.
//
// The one-warp-per-instance kernels keep a whole universe in registers (64 words per lane at
// 256 x 256: 255 registers, 8 warps per SM, issue/latency bound) and the four-warp kernel
// (quad.cuh) pays three block barriers per instance.  Here the unit of work is a strip: rows
// [q*32R, (q+1)*32R) of one instance, advanced by ONE warp with no cross-warp communication.
// Lane L holds rows L*R .. L*R+R-1 of the strip (R*WPL words); the rows just above and below the
// strip (toroidal) ride along as halo rows.  A persistent warp walks strips; per strip one
// elected lane issues TMA bulk copies (cp.async.bulk, mbarrier completion) of
//   - the strip's packed rows + the two halo rows, and
//   - only the rows of the caller's unpacked float32 / uint8 action that touch them,
// into the warp's private shared-memory slot.  The warp ballots the action rows into bit masks,
// drains the slot into registers, immediately re-issues the copy for its next strip, then XORs
// the action, advances one generation, emits the SpeedDetector partial sums and stores.
//
// Halo trick: lane 31's last-row triple is needed by nobody inside the strip, so lane 31 feeds
// the ABOVE-halo row through that slot of the rotate-up shuffle (lane 0 receives it as its
// upper neighbour); symmetrically lane 0 feeds the BELOW-halo row through the rotate-down
// shuffle.  Cost: one extra row triple + 2*WPL selects per strip.
//
// The U = WPL/R partial sums of an instance meet in two handle-owned 64-bit accumulators that
// also count arrivals; the last strip to arrive writes the result (no fence, no zero-fill
// launch: the accumulators re-zero themselves).  Geometry is compile-time (centred window, carle/env.py:119-132).
//
// Programmatic dependent launch: the kernel signals launch_dependents at once and waits on the
// previous grid (griddepcontrol.wait) before touching global memory, so back-to-back steps
// overlap launch latency and the prologue with the previous step's tail.
#pragma once
#include "kernels.cuh"

namespace carle {

// CUtensorMap as an opaque kernel parameter (the host encodes it with cuTensorMapEncodeTiled)
struct alignas(64) TensorMap { unsigned long long opaque[16]; };

namespace tma {
__device__ __forceinline__ void tensor_g2s_2d(uint32_t dst, const TensorMap* map, int c0, int c1,
                                              uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
}  // namespace tma

template <int WPL_, int R_, int AWIN_, typename T>
struct StripLayout {
    static constexpr int WPL = WPL_, R = R_, AWIN = AWIN_;
    static constexpr int H = 32 * WPL, U = WPL / R, ROWS = 32 * R;
    static constexpr int ROW0 = (H - AWIN) / 2;
    static constexpr int AW0 = ROW0 / 32, BIT0 = ROW0 % 32, C = AWIN / 32;
    static_assert(WPL % R == 0 && U >= 2, "at least two strips per instance");
    static_assert((WPL * 4) % 16 == 0, "bulk copies move whole 16-byte rows");
    static_assert(AWIN % 32 == 0 && AWIN < H && (H - AWIN) % 2 == 0, "window geometry");
    static_assert(AW0 + C + (BIT0 ? 1 : 0) <= WPL, "window words inside the row");

    // action rows strip q needs: the window rows among [q*ROWS - 1, (q+1)*ROWS] (body + the two
    // halo rows).  The window never touches the grid edge (AWIN < H, centred), so the toroidal
    // halos of the first / last strip are outside it and the range is contiguous: ONE bulk copy.
    static __host__ __device__ constexpr int act_lo(int q) {          // first window row
        return (q * ROWS - 1 > ROW0 ? q * ROWS - 1 : ROW0) - ROW0;
    }
    static __host__ __device__ constexpr int act_hi(int q) {          // one past the last
        return ((q + 1) * ROWS + 1 < ROW0 + AWIN ? (q + 1) * ROWS + 1 : ROW0 + AWIN) - ROW0;
    }
    static __host__ __device__ constexpr int act_rows(int q) {
        return act_hi(q) > act_lo(q) ? act_hi(q) - act_lo(q) : 0;
    }
    static __host__ __device__ constexpr int act_rows_max() {
        int m = 0;
        for (int q = 0; q < U; ++q) m = act_rows(q) > m ? act_rows(q) : m;
        return m;
    }
    // packed actions ([AW][AWPR] grid-aligned words per instance): rows of AWPR * 4 bytes, so the
    // row range of a strip is widened to the 16-byte granularity of the bulk copy
    static constexpr bool PACKED = IsPackedWords<T>::value;
    static constexpr int AWPR = C + (BIT0 ? 1 : 0);
    static constexpr int ALIGN = PACKED ? (16 / (AWPR * 4) > 0 ? 16 / (AWPR * 4) : 1) : 1;
    static_assert(!PACKED || (AWPR * 4 * ALIGN) % 16 == 0, "packed action rows per 16 bytes");
    static_assert(AWIN % ALIGN == 0, "window rows");
    static __host__ __device__ constexpr int act_lo_al(int q) { return act_lo(q) / ALIGN * ALIGN; }
    static __host__ __device__ constexpr int act_rows_al(int q) {
        return act_rows(q) ? (act_hi(q) + ALIGN - 1) / ALIGN * ALIGN - act_lo_al(q) : 0;
    }
    static __host__ __device__ constexpr int act_rows_al_max() {
        int m = 0;
        for (int q = 0; q < U; ++q) m = act_rows_al(q) > m ? act_rows_al(q) : m;
        return m;
    }
    static constexpr int ACT_ROWS = act_rows_al_max();
    static __host__ __device__ constexpr bool every_strip_has_rows() {
        for (int q = 0; q < U; ++q) if (act_rows(q) == 0) return false;
        return true;
    }
    // no strip without window rows (256x256 in 128-row strips): the "peek" at the action of a
    // window-less strip, and the registers it holds across the strip, are compiled out
    static constexpr bool ALL_ROWS = every_strip_has_rows();
    // every strip sees the SAME number of window rows (256x256 in 128-row strips: 33 each): the
    // ballot loop then has a compile-time trip count and is unrolled completely, which lets the
    // compiler run the shared-memory loads of later rows ahead of the ballots of earlier ones
    static __host__ __device__ constexpr bool same_rows_everywhere() {
        for (int q = 1; q < U; ++q) if (act_rows_al(q) != act_rows_al(0)) return false;
        return true;
    }
    static constexpr bool UNIFORM_ROWS = ALL_ROWS && same_rows_everywhere();
    static constexpr int ROW_BYTES = WPL * 4;
    // SWZ: the strip body arrives through a 2-D tensor-map copy with the 128-byte swizzle (16-byte
    // chunk index ^= 128-byte line index & 7), so the lanes' LDS.128 reads -- whose 128-byte lane
    // stride would otherwise put all 32 lanes on the same banks -- are conflict free.  Slot =
    // [body | halo above | halo below | action]; without SWZ [halo above | body | halo below | action].
    static constexpr bool SWZ = (R * WPL * 4 == 128);
    static constexpr int BODY_BYTES = ROWS * ROW_BYTES;
    static constexpr int BODY_LINES = BODY_BYTES / 128;             // 128-byte lines per strip
    static constexpr int STATE_BYTES = (ROWS + 2) * ROW_BYTES;
    static constexpr int ACT_ROW_BYTES = PACKED ? AWPR * 4 : AWIN * (int)sizeof(T);
    static constexpr int ACT_BYTES = ACT_ROWS * ACT_ROW_BYTES;
    static constexpr int SLOT_BYTES = STATE_BYTES + ACT_BYTES;      // multiple of 16
    static constexpr int MASK_BYTES = PACKED ? 0 : (ACT_ROWS * C * 4 + 15) / 16 * 16;
    static_assert(SLOT_BYTES % 16 == 0 && (ACT_ROW_BYTES * ALIGN) % 16 == 0, "bulk copy alignment");
    static_assert(PACKED || (4 * C) % 4 == 0, "four action rows = whole 16-byte groups of ballot masks");
    // (the 256x256 warp area is 13 KiB: one KiB more per warp and three CTAs no longer fit the 164 KiB
    //  shared-memory carve-out -- measured 98 -> 136 us)
    static constexpr int KEEP_BYTES = 0;
    static constexpr int keep_offset(int depth) { return (depth * SLOT_BYTES + MASK_BYTES + 8 * depth + 7) / 8 * 8; }
    static constexpr int warp_bytes(int depth) {
        return (keep_offset(depth) + KEEP_BYTES + (SWZ ? 1023 : 127)) / (SWZ ? 1024 : 128) * (SWZ ? 1024 : 128);
    }
};

// resident CTAs (4 warps each) per SM asked of ptxas
#ifndef CARLE_STRIP_INGEST_ROWS
#define CARLE_STRIP_INGEST_ROWS 4     // action rows whose loads are in flight at once in the ballot loop
#endif
#ifndef CARLE_STRIP_UNROLL_INGEST
#define CARLE_STRIP_UNROLL_INGEST 1   // compile-time trip count of the ballot loop where every strip has the same rows
#endif
#ifndef CARLE_STRIP_WARPS
#define CARLE_STRIP_WARPS 4           // warps per CTA (every warp works alone: CTA size only sets the occupancy grain)
#endif
#ifndef CARLE_STRIP_WARPS32
#define CARLE_STRIP_WARPS32 12        // 128-row strips of 256x256: resident warps per SM.  168 registers, 12 warps per SM (measured
                                      // faster than 4 CTAs x 128 registers: 83.5 vs 88.2 us, r1d_ab_strip_swizzle.txt)
#endif
// (the three targets are 28 / 16 / CARLE_STRIP_WARPS32 resident WARPS per SM)
constexpr int strip_min_ctas(int wpl, int r) {
    return (r * wpl <= 8 ? 28 : (r * wpl <= 16 ? 16 : CARLE_STRIP_WARPS32)) / CARLE_STRIP_WARPS;
}
template <int V> struct IntTag { static constexpr int value = V; };

// one generation of the strip held by this warp; `h` = the halo row this lane feeds into the
// shuffles (lane 31: the row above the strip, lane 0: the row below; unused elsewhere)
template <int R, int WPL, class Rule>
__device__ __forceinline__ void strip_generation(uint32_t (&x)[R][WPL], const uint32_t (&h)[WPL],
                                                 const Rule& rule, int lane) {
    static_assert(R >= 2, "strip rows per lane");
    const int up_lane = (lane + 31) & 31, dn_lane = (lane + 1) & 31;
    ca::Triple prev[WPL], cur[WPL], last[WPL], dn[WPL];
#pragma unroll
    for (int w = 0; w < WPL; ++w) {
        const int wl = (w + WPL - 1) % WPL, wr = (w + 1) % WPL;
        cur[w] = ca::row_triple(ca::west(x[0][wl], x[0][w]), x[0][w], ca::east(x[0][w], x[0][wr]));
        last[w] = ca::row_triple(ca::west(x[R - 1][wl], x[R - 1][w]), x[R - 1][w],
                                 ca::east(x[R - 1][w], x[R - 1][wr]));
        const ca::Triple th = ca::row_triple(ca::west(h[wl], h[w]), h[w], ca::east(h[w], h[wr]));
        // lane 31 sends the above-halo up the ring (to lane 0), lane 0 the below-halo down it
        const uint32_t ulo = (lane == 31) ? th.lo : last[w].lo, uhi = (lane == 31) ? th.hi : last[w].hi;
        const uint32_t dlo = (lane == 0) ? th.lo : cur[w].lo, dhi = (lane == 0) ? th.hi : cur[w].hi;
        prev[w].lo = __shfl_sync(0xFFFFFFFFu, ulo, up_lane);
        prev[w].hi = __shfl_sync(0xFFFFFFFFu, uhi, up_lane);
        dn[w].lo = __shfl_sync(0xFFFFFFFFu, dlo, dn_lane);
        dn[w].hi = __shfl_sync(0xFFFFFFFFu, dhi, dn_lane);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        ca::Triple nxt[WPL];
#pragma unroll
        for (int w = 0; w < WPL; ++w) {
            if (r + 2 < R) {
                const int wl = (w + WPL - 1) % WPL, wr = (w + 1) % WPL;
                nxt[w] = ca::row_triple(ca::west(x[r + 1][wl], x[r + 1][w]), x[r + 1][w],
                                        ca::east(x[r + 1][w], x[r + 1][wr]));
            } else if (r + 2 == R) {
                nxt[w] = last[w];
            } else {
                nxt[w] = dn[w];
            }
        }
#pragma unroll
        for (int w = 0; w < WPL; ++w) {
            x[r][w] = rule.from_triples(x[r][w], prev[w], cur[w], nxt[w]);
            prev[w] = cur[w];
            cur[w] = nxt[w];
        }
    }
}

template <int WPL, int R, int AWIN, class Rule, typename T, int DEPTH>
__global__ void __launch_bounds__(32 * CARLE_STRIP_WARPS, strip_min_ctas(WPL, R))
step_strip_kernel(const __grid_constant__ StepParams p, const __grid_constant__ TensorMap state_map) {
    using L = StripLayout<WPL, R, AWIN, T>;
    constexpr int H = L::H, U = L::U, ROWS = L::ROWS, C = L::C, ROW0 = L::ROW0;
    constexpr int WORDS = R * WPL;
    static_assert(WORDS % 4 == 0, "vector loads");
    extern __shared__ __align__(1024) unsigned char strip_smem[];
    __shared__ unsigned int s_done;
    __shared__ double s_sd;
    const int lane = threadIdx.x & 31;
    // warp index through a shuffle: tells the compiler it is warp-uniform, so the strip
    // bookkeeping and the bulk-copy operands live in uniform registers (no per-copy broadcast)
    const int wib = __shfl_sync(0xFFFFFFFFu, (int)(threadIdx.x >> 5), 0);
    const int warps_per_block = blockDim.x >> 5;
    // (unit indices are 32-bit: the launcher refuses batches of 2^31 strips or more)
    const int nwarps = (int)gridDim.x * warps_per_block;
    // block-interleaved rank: a partial last trip is spread evenly over the CTAs (and SMs);
    // blocked rank: the four warps of a CTA stream the four strips of one instance
    const int rank = p.rank_blocked ? (int)blockIdx.x * warps_per_block + wib
                                    : wib * (int)gridDim.x + (int)blockIdx.x;
    const int rot = p.rank_blocked ? 0 : wib;
    unsigned char* wbase = strip_smem + (size_t)wib * L::warp_bytes(DEPTH);
    uint32_t* amask = reinterpret_cast<uint32_t*>(wbase + DEPTH * L::SLOT_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(wbase + DEPTH * L::SLOT_BYTES + L::MASK_BYTES);

    pdl_launch_dependents();
    if (threadIdx.x == 0) { s_done = 0u; s_sd = 0.0; }
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < DEPTH; ++s) tma::mbar_init(bars + s, 1);
        tma::fence_mbar_init();
    }
    __syncthreads();
    pdl_wait();                                   // the previous step's state / flags are final

    const Rule rule(p);
    const char* in_bytes = reinterpret_cast<const char*>(p.in);
    const char* act_bytes = static_cast<const char*>(p.raw);
    const long long act_stride = p.raw_inst_stride * (long long)sizeof(T);
    const int total = (int)(p.n * U);
    // unit `v` of this warp's walk -> strip index (see StepParams::reverse; `total` is a multiple of
    // U, so the U strips of an instance stay on U neighbouring ranks of one trip)
    auto unit_strip = [&](int v) { return p.reverse ? total - 1 - v : v; };

    // strip `u` of trip `trip`: the U strips of an instance sit on U consecutive ranks of one
    // trip (the host keeps gridDim.x a multiple of U); the rotation by wib + trip mixes strips
    // with and without window rows inside every CTA
    auto strip_q = [&](int u, int trip) { return (u + rot + trip) & (U - 1); };

    // called by the whole (converged) warp with warp-uniform arguments; one elected lane issues
    auto issue = [&](int s, int u, int trip, uint32_t dep) {   // dep == 0 (StepParams::zero)
        const long long inst = u / U;
        const int q = strip_q(u, trip), r0 = q * ROWS;
        const uint32_t slot = tma::smem_u32(wbase + s * L::SLOT_BYTES);
        const uint32_t bar = tma::smem_u32(bars + s);
        const int a_lo = L::act_lo_al(q), a_n = L::act_rows_al(q);
        const char* src = in_bytes + inst * (long long)(H * L::ROW_BYTES);
        const char* asrc = act_bytes + inst * act_stride + a_lo * L::ACT_ROW_BYTES;
        // state: rows r0-1 .. r0+ROWS as one copy, or two when a halo wraps around the torus
        const char* src0 = src + (q == 0 ? (H - 1) : (r0 - 1)) * L::ROW_BYTES;
        const uint32_t bytes0 = (q == 0) ? L::ROW_BYTES
                              : (q == U - 1) ? (ROWS + 1) * L::ROW_BYTES : (ROWS + 2) * L::ROW_BYTES;
        const bool two = (q == 0) || (q == U - 1);
        const uint32_t dst1 = slot + (q == 0 ? L::ROW_BYTES : (ROWS + 1) * L::ROW_BYTES);
        const uint32_t bytes1 = (q == 0) ? (ROWS + 1) * L::ROW_BYTES : L::ROW_BYTES;
        if (tma::elect_one()) {
            tma::mbar_expect_tx_u32(bar, L::STATE_BYTES + (uint32_t)a_n * L::ACT_ROW_BYTES + dep);
            if constexpr (L::SWZ) {
                // body: one swizzled tensor copy of BODY_LINES 128-byte lines; halos: two rows
                const long long line0 = (inst * H + r0) * (long long)L::ROW_BYTES / 128;
                tma::tensor_g2s_2d(slot, &state_map, 0, (int)line0, bar);
                tma::bulk_g2s_u32(slot + L::BODY_BYTES, src + ((r0 + H - 1) & (H - 1)) * L::ROW_BYTES,
                                  L::ROW_BYTES, bar);
                tma::bulk_g2s_u32(slot + L::BODY_BYTES + L::ROW_BYTES,
                                  src + ((r0 + ROWS) & (H - 1)) * L::ROW_BYTES, L::ROW_BYTES, bar);
            } else {
                tma::bulk_g2s_u32(slot, src0, bytes0, bar);
                if (two) tma::bulk_g2s_u32(dst1, src, bytes1, bar);
            }
            if (a_n > 0)    // (the policy is re-created here rather than held in two registers across the loop)
                tma::bulk_g2s_hint(slot + L::STATE_BYTES, asrc, (uint32_t)a_n * L::ACT_ROW_BYTES, bar,
                                   tma::l2_policy(p.act_evict_first != 0));
        }
        __syncwarp();
    };

#pragma unroll
    for (int s = 0; s < DEPTH; ++s) {
        const int v = rank + s * nwarps;
        if (v < total) issue(s, unit_strip(v), s, 0u);
    }

    bool warp_not_one = false, warp_any = false, warp_nonbin = false;
    // Hand-over of the fused sums WITHOUT waiting on an atomic.  The U strip partials of an
    // instance meet in two handle-owned 64-bit accumulators (zero between launches)
    //   acc[0] = live | sh << 20 | arrivals << 56,   acc[1] = window-live | sw << 20 | arrivals << 56
    // Every strip adds its partials with fire-and-forget reductions (red.add) and moves on: the loop
    // below carries NO state of the hand-over.  BEHIND the loop the warp of an instance's first strip
    // (rank % U == 0; its partners run on the neighbouring ranks of the same trip) collects its
    // instances, one per LANE: it reads the two words (polling in the rare case that a partner's add
    // of the last trips has not landed yet), writes the sums, re-zeroes the words and -- with the
    // SpeedDetector tail fused in (p.sd_com) -- turns the sums into centre of mass, velocity and the
    // warp's share of the sum of squared velocities.  An atomic with a return value instead (the strip
    // whose add returns U-1 owns the sum) made every warp wait out the L2 round trip behind the atomic
    // (15 % of the stall samples, profiles/r1d_step_strip_kernel_cfg3.summary.txt); reading the words
    // back inside the loop one trip later (rounds 1e - 2f) cost six loop-carried registers and ~3 us
    // of 95 (profiles/r2g_ab_sums_handover.txt).
    const bool sd_on = CARLE_FEAT_SD && p.sd_com != nullptr;
    const bool sd_settler = (rank & (U - 1)) == 0;
    bool warp_may_reset = false;                    // some strip of this warp fenced (all ones / non-binary)
    int trip = 0;
    for (int v = rank; v < total; v += nwarps, ++trip) {
        const int u = unit_strip(v);
        const int s = trip % DEPTH;
        const int inst32 = u / U;
        const long long inst = inst32;
        const int q = strip_q(u, trip), r0 = q * ROWS;
        const int a_lo = L::act_lo_al(q), act_rows = L::act_rows_al(q);
        const unsigned char* slot = wbase + s * L::SLOT_BYTES;
        // A strip without window rows never sees the action, yet the batch-wide master reset
        // (retire_fused) needs its stores fenced when every toggle of the batch is 1.0: peek at the
        // first 16 bytes of the instance's action -- unless they are all ones no reset can fire.
        uint4 peek = make_uint4(0u, 0u, 0u, 0u);
        if constexpr (!L::ALL_ROWS)
            if (act_rows == 0) peek = __ldg(reinterpret_cast<const uint4*>(act_bytes + inst * act_stride));
        tma::mbar_wait(bars + s, (uint32_t)((trip / DEPTH) & 1));

        // ---- action rows -> ballot masks (carle/env.py:179-182, 191) ----
        const T* a = reinterpret_cast<const T*>(slot + L::STATE_BYTES) + lane;
        NonBinary nb;                                     // some toggle is neither 0 nor 1 (kernels.cuh)
        if constexpr (!L::PACKED) {
            int j = 0;
            // CARLE_STRIP_INGEST_ROWS rows (x C loads) in flight per trip
            auto rows_at_once = [&](auto rows_tag) {
                constexpr int NR = decltype(rows_tag)::value;
                for (; j + NR <= act_rows; j += NR) {
                    T v[NR][C];
#pragma unroll
                    for (int i = 0; i < NR; ++i)
#pragma unroll
                        for (int c = 0; c < C; ++c) v[i][c] = a[(j + i) * AWIN + c * 32];
                    uint32_t m[NR * C];
#pragma unroll
                    for (int i = 0; i < NR; ++i)
#pragma unroll
                        for (int c = 0; c < C; ++c) m[i * C + c] = __ballot_sync(0xFFFFFFFFu, v[i][c] != T(0));
                    // (the masks of these rows are contiguous and 16-byte aligned: j, NR * C are multiples
                    //  of 4 and the mask area starts on a 16-byte boundary -- one vector store per four)
                    if (lane == 0) {
#pragma unroll
                        for (int k = 0; k < NR * C; k += 4)
                            *reinterpret_cast<uint4*>(amask + j * C + k) = make_uint4(m[k], m[k + 1], m[k + 2], m[k + 3]);
                    }
                    nb.see_all(reinterpret_cast<const T(&)[NR * C]>(v));
                }
            };
            if constexpr (L::UNIFORM_ROWS && CARLE_STRIP_UNROLL_INGEST) {
                constexpr int AR = L::act_rows_al(0), FULL = AR / 4 * 4;
#pragma unroll
                for (int jj = 0; jj < FULL; jj += 4) {
                    T v[4][C];
                    uint32_t m[4 * C];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int c = 0; c < C; ++c) v[i][c] = a[(jj + i) * AWIN + c * 32];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int c = 0; c < C; ++c) m[i * C + c] = __ballot_sync(0xFFFFFFFFu, v[i][c] != T(0));
                    if (lane == 0) {
#pragma unroll
                        for (int k = 0; k < 4 * C; k += 4)
                            *reinterpret_cast<uint4*>(amask + jj * C + k) = make_uint4(m[k], m[k + 1], m[k + 2], m[k + 3]);
                    }
                    nb.see_all(reinterpret_cast<const T(&)[4 * C]>(v));
                }
#pragma unroll
                for (int jj = FULL; jj < AR; ++jj)
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const T v = a[jj * AWIN + c * 32];
                        const uint32_t m = __ballot_sync(0xFFFFFFFFu, v != T(0));
                        nb.see(v);
                        if (lane == 0) amask[jj * C + c] = m;
                    }
            } else {
            if constexpr (CARLE_STRIP_INGEST_ROWS > 4) rows_at_once(IntTag<CARLE_STRIP_INGEST_ROWS>{});
            rows_at_once(IntTag<4>{});
            for (; j < act_rows; ++j) {                  // tail rows (never read past the slot)
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const T v = a[j * AWIN + c * 32];
                    const uint32_t m = __ballot_sync(0xFFFFFFFFu, v != T(0));
                    nb.see(v);
                    if (lane == 0) amask[j * C + c] = m;
                }
            }
            }
        }
        const bool inst_nonbin = __any_sync(0xFFFFFFFFu, nb.any_lane());
        warp_nonbin |= inst_nonbin;
        __syncwarp();

        // ---- drain the slot: this lane's rows, its halo row, its action masks ----
        uint32_t x[R][WPL], h[WPL];
        {
            // (SWZ: chunk c of this lane's 128-byte line sits at chunk position c ^ (line & 7))
            constexpr int LANE_BYTES = WORDS * 4;
            const int line = lane * LANE_BYTES / 128, chunk0 = (lane * LANE_BYTES % 128) / 16;
            const unsigned char* body = L::SWZ ? slot + line * 128 : slot + (1 + lane * R) * L::ROW_BYTES;
#pragma unroll
            for (int i = 0; i < WORDS / 4; ++i) {
                const uint4 t = *reinterpret_cast<const uint4*>(
                    L::SWZ ? body + (((chunk0 + i) ^ (line & 7)) << 4) : body + 16 * i);
                (&x[0][0])[4 * i + 0] = t.x; (&x[0][0])[4 * i + 1] = t.y;
                (&x[0][0])[4 * i + 2] = t.z; (&x[0][0])[4 * i + 3] = t.w;
            }
            const uint4* hs = reinterpret_cast<const uint4*>(
                L::SWZ ? slot + L::BODY_BYTES + ((lane == 0) ? L::ROW_BYTES : 0)
                       : slot + ((lane == 0) ? (ROWS + 1) : 0) * L::ROW_BYTES);
#pragma unroll
            for (int i = 0; i < WPL / 4; ++i) {
                const uint4 t = hs[i];
                h[4 * i + 0] = t.x; h[4 * i + 1] = t.y; h[4 * i + 2] = t.z; h[4 * i + 3] = t.w;
            }
        }
        uint32_t am[R][C], hm[C];
        uint32_t pw[R][L::AWPR], hw[L::AWPR];             // packed actions: grid-aligned toggle words
        bool inst_not_one;
        if constexpr (L::PACKED) {
            // the action arrives packed: every lane fetches the words of its own rows (and of its
            // halo row); the batch-wide flags come from the words themselves
            const uint32_t* aw = reinterpret_cast<const uint32_t*>(slot + L::STATE_BYTES);
            uint32_t seen = 0u, all_set = 0xFFFFFFFFu;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int idx = r0 + lane * R + r - ROW0 - a_lo;
                const bool in = (unsigned)idx < (unsigned)act_rows;
#pragma unroll
                for (int j = 0; j < L::AWPR; ++j) {
                    constexpr int CC0 = ROW0;
                    const uint32_t valid = packed_valid_mask<CC0, AWIN>(j);
                    const uint32_t w = in ? aw[(in ? idx : 0) * L::AWPR + j] : 0u;
                    pw[r][j] = w & valid;
                    seen |= w & valid;
                    all_set &= in ? (w | ~valid) : 0xFFFFFFFFu;
                }
            }
            {
                const int idx = ((lane == 0) ? r0 + ROWS : r0 - 1) - ROW0 - a_lo;
                const bool in = (lane == 0 || lane == 31) && (unsigned)idx < (unsigned)act_rows;
#pragma unroll
                for (int j = 0; j < L::AWPR; ++j) {
                    constexpr int CC0 = ROW0;
                    hw[j] = in ? (aw[(in ? idx : 0) * L::AWPR + j] & packed_valid_mask<CC0, AWIN>(j)) : 0u;
                }
            }
            warp_any |= __any_sync(0xFFFFFFFFu, seen != 0u);
            const bool seen_not_one = __any_sync(0xFFFFFFFFu, all_set != 0xFFFFFFFFu);
            bool peek_not_one = false;                     // window-less strip: first 16 bytes of the action
            {
                const uint32_t pk[4] = {peek.x, peek.y, peek.z, peek.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    constexpr int CC0 = ROW0;
                    const uint32_t valid = packed_valid_mask<CC0, AWIN>(k % L::AWPR);
                    peek_not_one |= (pk[k] & valid) != valid;
                }
            }
            inst_not_one = (L::ALL_ROWS || act_rows) ? seen_not_one : peek_not_one;
            warp_not_one |= (L::ALL_ROWS || act_rows) ? seen_not_one : false;
        } else {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int idx = r0 + lane * R + r - ROW0 - a_lo;         // slot row of x[r]'s action row
            const bool in = (unsigned)idx < (unsigned)act_rows;
            if constexpr (C == 2) {
                const uint2 t = *reinterpret_cast<const uint2*>(amask + (in ? idx : 0) * C);
                am[r][0] = in ? t.x : 0u;
                am[r][1] = in ? t.y : 0u;
            } else {
#pragma unroll
                for (int c = 0; c < C; ++c) am[r][c] = in ? amask[(in ? idx : 0) * C + c] : 0u;
            }
        }
        {
            // lane 31 carries the row above the strip, lane 0 the row below (both outside the
            // window whenever they wrap around the torus)
            const int idx = ((lane == 0) ? r0 + ROWS : r0 - 1) - ROW0 - a_lo;
            const bool in = (lane == 0 || lane == 31) && (unsigned)idx < (unsigned)act_rows;
            if constexpr (C == 2) {
                const uint2 t = *reinterpret_cast<const uint2*>(amask + (in ? idx : 0) * C);
                hm[0] = in ? t.x : 0u;
                hm[1] = in ? t.y : 0u;
            } else {
#pragma unroll
                for (int c = 0; c < C; ++c) hm[c] = in ? amask[(in ? idx : 0) * C + c] : 0u;
            }
        }
        // batch-wide flags: some toggle != 0, and some toggle != 1.0 (master reset, env.py:208).  A
        // zero toggle settles the second, and the masks already say whether there is one; only if
        // EVERY toggle of the strip's rows is non-zero are the values themselves compared with 1.0.
        {
            uint32_t seen = 0u, all_set = 0xFFFFFFFFu;
            for (int k = lane; k < act_rows * C; k += 32) {
                const uint32_t m = amask[k];
                seen |= m;
                all_set &= m;
            }
            warp_any |= __any_sync(0xFFFFFFFFu, seen != 0u);
            bool seen_not_one = __any_sync(0xFFFFFFFFu, all_set != 0xFFFFFFFFu);
            if (!seen_not_one && act_rows) {             // rare; the slot is not refilled yet
                uint32_t differs = 0u;
                for (int k = 0; k < act_rows; ++k)
#pragma unroll
                    for (int c = 0; c < C; ++c)
                        differs |= bits_of(a[k * AWIN + c * 32]) ^ OneBits<T>::value;
                seen_not_one = __any_sync(0xFFFFFFFFu, differs != 0u);
            }
            constexpr uint32_t ONES = sizeof(T) == 1 ? 0x01010101u : OneBits<T>::value;
            inst_not_one = (L::ALL_ROWS || act_rows) ? seen_not_one
                                    : (peek.x != ONES || peek.y != ONES || peek.z != ONES || peek.w != ONES);
            warp_not_one |= seen_not_one;
        }
        }
        // The refill below overwrites the slot through the async proxy, and a bank-conflicted LDS
        // can still be queued in the LSU when later instructions issue: the refill's byte count
        // depends on one register of every load above (dep == 0, see StepParams::zero).
        uint32_t dep = 0u;
#pragma unroll
        for (int i = 0; i < WORDS / 4; ++i) dep ^= (&x[0][0])[4 * i];
#pragma unroll
        for (int i = 0; i < WPL / 4; ++i) dep ^= h[4 * i];
        if constexpr (L::PACKED) {
#pragma unroll
            for (int r = 0; r < R; ++r) dep ^= pw[r][0] ^ pw[r][L::AWPR - 1];
            dep ^= hw[0] ^ hw[L::AWPR - 1];
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r) dep ^= am[r][0];
            dep ^= hm[0];
        }
        dep &= p.zero;
        __syncwarp();                                   // slot and masks are drained: refill
        {
            const int nv = v + DEPTH * nwarps;
            if (nv < total) issue(s, unit_strip(nv), trip + DEPTH, dep);
        }

        if constexpr (L::PACKED) {
#pragma unroll
            for (int j = 0; j < L::AWPR; ++j) {
#pragma unroll
                for (int r = 0; r < R; ++r) x[r][L::AW0 + j] ^= pw[r][j];
                h[L::AW0 + j] ^= hw[j];
            }
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r) xor_action_row<WPL, L::AW0, L::BIT0, C>(x[r], am[r]);
            xor_action_row<WPL, L::AW0, L::BIT0, C>(h, hm);
        }
        strip_generation<R, WPL>(x, h, rule, lane);

        // ---- fused SpeedDetector sums (carle/mcl.py:773-779) ----
        if (p.red) {
            uint32_t live = 0, sh = 0, sw = 0, wl = 0;
            ca::strip_lane_sums<WPL, R, AWIN>(x, r0 + lane * R, live, sh, sw, wl);
            // (a strip holds at most 2^15 cells: live and window-live share one reduction)
            static_assert(ROWS * 32 * WPL <= (1 << 15), "packed live / window-live reduction");
            const uint32_t lw = __reduce_add_sync(0xFFFFFFFFu, live | (wl << 16));
            live = lw & 0xFFFFu;
            wl = lw >> 16;
            sh = __reduce_add_sync(0xFFFFFFFFu, sh);
            sw = __reduce_add_sync(0xFFFFFFFFu, sw);
            if (lane == 0) {
                unsigned long long* acc = reinterpret_cast<unsigned long long*>(p.strip_part) + inst * 2;
                const unsigned long long add_a = live | ((unsigned long long)sh << 20) | (1ull << 56);
                const unsigned long long add_b = wl | ((unsigned long long)sw << 20) | (1ull << 56);
                asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" :: "l"(acc), "l"(add_a) : "memory");
                asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" :: "l"(acc + 1), "l"(add_b) : "memory");
            }
        }
        // ---- next state: R*WPL contiguous words per lane ----
        {
            uint4* dst = reinterpret_cast<uint4*>(p.out + inst * (long long)(H * WPL) +
                                                  (long long)(r0 + lane * R) * WPL);
#pragma unroll
            for (int i = 0; i < WORDS / 4; ++i)
                dst[i] = make_uint4((&x[0][0])[4 * i + 0], (&x[0][0])[4 * i + 1],
                                    (&x[0][0])[4 * i + 2], (&x[0][0])[4 * i + 3]);
        }
        if (p.reward_zero && !sd_on && q == 0 && lane == 0) p.reward_zero[inst] = 0.f;
        if (CARLE_FEAT_OBS && p.obs) emit_obs_any<WORDS>(p, &x[0][0], (inst * H + r0) * (long long)(32 * WPL), lane);
        fence_if_all_ones(inst_not_one && !inst_nonbin);
        warp_may_reset |= !(inst_not_one && !inst_nonbin);
    }
    double sd_local = 0.0;
    bool sd_primed = false;
    if (p.red && sd_settler) {
        // ---- the sums of this warp's instances, one instance per lane (see above) ----
        const int primed_raw = sd_on ? *p.sd_primed : 0;        // (set by the previous step; looked at behind the poll)
        constexpr unsigned long long F20 = (1ull << 20) - 1, F36 = (1ull << 36) - 1;
        for (int t0 = 0; t0 < trip; t0 += 32) {
            const int t = t0 + lane;
            if (t < trip) {
                const long long inst = unit_strip(rank + t * nwarps) / U;
                unsigned long long* acc = reinterpret_cast<unsigned long long*>(p.strip_part) + inst * 2;
                // (the previous centres of mass are requested ahead of the poll: one memory round trip
                //  for everything instead of three in a row at the very end of the kernel)
                float prev_h = 0.f, prev_w = 0.f;
                if (sd_on) { prev_h = p.sd_com_prev[inst]; prev_w = p.sd_com_prev[p.n + inst]; }
                unsigned long long a, b;
                do {
                    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(a) : "l"(acc) : "memory");
                    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(b) : "l"(acc + 1) : "memory");
                } while ((a >> 56) != (unsigned long long)U || (b >> 56) != (unsigned long long)U);
                acc[0] = 0ull;
                acc[1] = 0ull;
                sd_primed = primed_raw != 0;
                const uint32_t live = (uint32_t)(a & F20);
                const unsigned long long sh = (a >> 20) & F36, sw = (b >> 20) & F36;
                longlong2* o = reinterpret_cast<longlong2*>(p.red + inst * 4);
                o[0] = make_longlong2((long long)live, (long long)sh);
                o[1] = make_longlong2((long long)sw, (long long)(b & F20));
                if (sd_on)
                    sd_local += speed_instance(p, inst, sd_primed, prev_h, prev_w, live, sh, sw);
            }
        }
        if (sd_on) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) sd_local += __shfl_xor_sync(0xFFFFFFFFu, sd_local, off);
        }
        // (the sums must be ordered ahead of a master reset's clear like the state stores are)
        sd_primed = primed_raw != 0;
        if (warp_may_reset) __threadfence();
    } else if (sd_on) {
        sd_primed = *p.sd_primed != 0;
    }
    // ---- retirement: warp -> block (shared memory) -> grid (global), flags inside the atomics ----
    if (sd_on) speed_warp_done(&s_sd, lane, sd_local);
    const int last_of_grid = retire_fused<T>(p, &s_done, lane, warps_per_block, warp_not_one,
                                             warp_any, warp_nonbin, sd_on ? &s_sd : nullptr);
    if (last_of_grid == 2) clear_after_reset(p, lane);
    if (last_of_grid && sd_on) speed_grid_done(p, lane, sd_primed, last_of_grid == 2);
}

}  // namespace carle
