// ca_core.cuh — bit-sliced Life-like cellular-automaton arithmetic on 32-cell words.
//
// Replaces the reference's  conv2d(3x3 Moore, circular) -> `elem == count` ->
// (1-u)*birth + u*survive  pipeline (carle/env.py:219-229) with Boolean algebra on
// bit planes: each 32-bit word holds 32 horizontally adjacent cells, so one LOP3
// advances 32 cells.  Pure functions on uint32 so the same text compiles for the
// device (kernels) and for the host (tests/cpu_twin, used only by the tests).
#pragma once
#if defined(__CUDACC_RTC__)
// NVRTC (run-time rule specialisation, carle_abi.cu): no host headers
typedef unsigned char uint8_t;
typedef unsigned short uint16_t;
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
#else
#include <stdint.h>
#endif

#if defined(__CUDACC__)
#define CA_HD __host__ __device__ __forceinline__
#else
#define CA_HD inline
#endif

namespace ca {

// ---- horizontal neighbours ------------------------------------------------------
// west(x)[b] = cell at column-1, east(x)[b] = cell at column+1; `prev`/`next` are the
// words to the left/right in the same row (toroidal wrap is the caller's indexing).
CA_HD uint32_t west(uint32_t prev, uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(prev, x, 1);
#else
    return (x << 1) | (prev >> 31);
#endif
}
CA_HD uint32_t east(uint32_t x, uint32_t next) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(x, next, 1);
#else
    return (x >> 1) | (next << 31);
#endif
}

// ---- LOP3: any 3-input Boolean function in one instruction ----------------------------
// LUT bit index = a<<2 | b<<1 | c (the PTX lop3.b32 convention).
template <uint32_t LUT>
CA_HD uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return r;
#else
    uint32_t r = 0;
    if (LUT & 0x01u) r |= ~a & ~b & ~c;
    if (LUT & 0x02u) r |= ~a & ~b & c;
    if (LUT & 0x04u) r |= ~a & b & ~c;
    if (LUT & 0x08u) r |= ~a & b & c;
    if (LUT & 0x10u) r |= a & ~b & ~c;
    if (LUT & 0x20u) r |= a & ~b & c;
    if (LUT & 0x40u) r |= a & b & ~c;
    if (LUT & 0x80u) r |= a & b & c;
    return r;
#endif
}
constexpr uint32_t LUT_XOR3 = 0x96, LUT_MAJ = 0xE8, LUT_MUX = 0xCA;  // MUX: a ? b : c

// ---- row triple: number of live cells among (west, self, east), 0..3, as 2 planes --
struct Triple { uint32_t lo, hi; };

CA_HD Triple row_triple(uint32_t w, uint32_t x, uint32_t e) {
    Triple t;
    t.lo = lop3<LUT_XOR3>(w, x, e);
    t.hi = lop3<LUT_MAJ>(w, x, e);
    return t;
}

// ---- 3x3 sum including the centre, 0..9, in carry-save form --------------------------
// sum9 = t0 + 2*(k0 + t1) + 4*k1
struct Sum9 { uint32_t t0, k0, t1, k1; };

CA_HD Sum9 add3(Triple a, Triple c, Triple b) {
    Sum9 s;
    s.t0 = lop3<LUT_XOR3>(a.lo, c.lo, b.lo);
    s.k0 = lop3<LUT_MAJ>(a.lo, c.lo, b.lo);
    s.t1 = lop3<LUT_XOR3>(a.hi, c.hi, b.hi);
    s.k1 = lop3<LUT_MAJ>(a.hi, c.hi, b.hi);
    return s;
}

// ---- rule tables ----------------------------------------------------------------------
// For a cell with state x and sum9 parity plane t0, the next state is a 3-input Boolean
// function of (k0, t1, k1) -- an 8-bit LOP3 truth table.  Index = k0<<2 | t1<<1 | k1,
// u = k0 + t1 + 2*k1 (0..4), sum9 = t0 + 2u; a dead cell with sum9 = n is born iff n in B,
// a live cell survives iff (sum9 - 1) in S.
CA_HD constexpr uint32_t class_lut(uint32_t birth, uint32_t survive, int x, int t0) {
    uint32_t lut = 0;
    for (int idx = 0; idx < 8; ++idx) {
        int k0 = (idx >> 2) & 1, t1 = (idx >> 1) & 1, k1 = idx & 1;
        int sum9 = t0 + 2 * (k0 + t1 + 2 * k1);
        bool on = false;
        if (x == 0) { if (sum9 <= 8) on = (birth >> sum9) & 1u; }
        else        { if (sum9 >= 1) on = (survive >> (sum9 - 1)) & 1u; }
        if (on) lut |= 1u << idx;
    }
    return lut;
}

// One half of the rule: h(x, k0, t1, k1) = x ? lut3<L1> : lut3<L0>, in 1..3 LOP3.
template <uint32_t L0, uint32_t L1>
CA_HD uint32_t half_rule(uint32_t x, uint32_t k0, uint32_t t1, uint32_t k1) {
    if constexpr (L0 == L1) {
        if constexpr (L0 == 0u) return 0u;
        else if constexpr (L0 == 0xFFu) return 0xFFFFFFFFu;
        else return lop3<L0>(k0, t1, k1);
    } else if constexpr (L0 == 0u) {
        if constexpr (L1 == 0xFFu) return x;
        else return x & lop3<L1>(k0, t1, k1);
    } else if constexpr (L1 == 0u) {
        if constexpr (L0 == 0xFFu) return ~x;
        else return ~x & lop3<L0>(k0, t1, k1);
    } else if constexpr (L0 == 0xFFu) {
        return ~x | lop3<L1>(k0, t1, k1);
    } else if constexpr (L1 == 0xFFu) {
        return x | lop3<L0>(k0, t1, k1);
    } else {
        return lop3<LUT_MUX>(x, lop3<L1>(k0, t1, k1), lop3<L0>(k0, t1, k1));
    }
}

// Rule known at compile time: next = t0 ? h1(x, ..) : h0(x, ..); at most 7 LOP3, 4 for
// Conway's Life (h1 does not depend on x, h0 = x & one table).
template <uint32_t BIRTH, uint32_t SURVIVE>
CA_HD uint32_t next_static(uint32_t x, Sum9 s) {
    constexpr uint32_t L00 = class_lut(BIRTH, SURVIVE, 0, 0);
    constexpr uint32_t L01 = class_lut(BIRTH, SURVIVE, 0, 1);
    constexpr uint32_t L10 = class_lut(BIRTH, SURVIVE, 1, 0);
    constexpr uint32_t L11 = class_lut(BIRTH, SURVIVE, 1, 1);
    const uint32_t h0 = half_rule<L00, L10>(x, s.k0, s.t1, s.k1);   // even sum9
    const uint32_t h1 = half_rule<L01, L11>(x, s.k0, s.t1, s.k1);   // odd sum9
    return lop3<LUT_MUX>(s.t0, h1, h0);
}

// Conway's Life straight from the three row triples in SEVEN LOP3 (add3 + next_static take 8;
// per word-generation: 2 funnel shifts + 2 for the row triple + 7 = 11 integer-pipe
// instructions, which is what bounds the register-resident multi-generation kernels).
// L = lo_a + lo_c + lo_b and H = hi_a + hi_c + hi_b (sum9 = L + 2H) are re-encoded as
//   A = [L in {1,2}], B = L & 1, U = [H in {1,2}], V = H & 1      (one LOP3 each)
// and next = f(x, A, B, U, V) is a three-LOP3 network.  Found by exhaustive search over all
// three-LOP3 networks on every 2-bit encoding of L and H (none exists on the binary encoding
// add3 produces, and no two-LOP3 network on any encoding); checked on all 2^7 inputs by
// tests/test_core_math_cpu.py.
CA_HD uint32_t life_from_triples(uint32_t x, Triple a, Triple c, Triple b) {
    const uint32_t A = lop3<0x7E>(a.lo, c.lo, b.lo), B = lop3<LUT_XOR3>(a.lo, c.lo, b.lo);
    const uint32_t U = lop3<0x7E>(a.hi, c.hi, b.hi), V = lop3<LUT_XOR3>(a.hi, c.hi, b.hi);
    const uint32_t y1 = lop3<0x27>(x, A, B);
    const uint32_t y2 = lop3<0x64>(B, U, y1);
    return lop3<0x82>(A, V, y2);
}

// Morley / "Move" (B368/S245, BASELINE config 3) and HighLife (B36/S23) on the same encoding in
// EIGHT LOP3 (add3 + next_static take 11 and 10): four for (A, B, U, V), then a four-LOP3 network
// found by the same kind of search (the last two nodes by functional decomposition of the rule
// over every pair of earlier signals); checked on all 2^7 inputs by tests/test_core_math_cpu.py.
CA_HD uint32_t morley_from_triples(uint32_t x, Triple a, Triple c, Triple b) {
    const uint32_t A = lop3<0x7E>(a.lo, c.lo, b.lo), B = lop3<LUT_XOR3>(a.lo, c.lo, b.lo);
    const uint32_t U = lop3<0x7E>(a.hi, c.hi, b.hi), V = lop3<LUT_XOR3>(a.hi, c.hi, b.hi);
    const uint32_t y1 = lop3<104>(x, A, V);
    const uint32_t y2 = lop3<20>(A, U, y1);
    const uint32_t y3 = lop3<28>(B, V, y1);
    return lop3<18>(U, y2, y3);
}
CA_HD uint32_t highlife_from_triples(uint32_t x, Triple a, Triple c, Triple b) {
    const uint32_t A = lop3<0x7E>(a.lo, c.lo, b.lo), B = lop3<LUT_XOR3>(a.lo, c.lo, b.lo);
    const uint32_t U = lop3<0x7E>(a.hi, c.hi, b.hi), V = lop3<LUT_XOR3>(a.hi, c.hi, b.hi);
    const uint32_t y1 = lop3<5>(x, A, B);
    const uint32_t y2 = lop3<35>(x, A, B);
    const uint32_t y3 = lop3<105>(A, V, y1);
    return lop3<40>(U, y2, y3);
}

// Day & Night (B3678/S34678) in EIGHT LOP3 as well, on the encoding B = [L >= 2], V = [H >= 2] (no
// four-LOP3 network exists for it on the parity encoding the other rules use): tools/lop3_search/final4.c
// with -DBIRTH=0x1C8 -DSURV=0x1D8, encoding 4; checked on all 2^7 inputs by tests/test_core_math_cpu.py.
CA_HD uint32_t daynight_from_triples(uint32_t x, Triple a, Triple c, Triple b) {
    const uint32_t A = lop3<0x7E>(a.lo, c.lo, b.lo), B = lop3<LUT_MAJ>(a.lo, c.lo, b.lo);
    const uint32_t U = lop3<0x7E>(a.hi, c.hi, b.hi), V = lop3<LUT_MAJ>(a.hi, c.hi, b.hi);
    const uint32_t y1 = lop3<74>(x, A, V);
    const uint32_t y2 = lop3<105>(x, B, U);
    const uint32_t y3 = lop3<148>(A, B, y1);
    return lop3<226>(V, y2, y3);
}

// compile-time rule from the row triples above / of / below the cell
template <uint32_t BIRTH, uint32_t SURVIVE>
CA_HD uint32_t next_static_triples(uint32_t x, Triple a, Triple c, Triple b) {
    if constexpr (BIRTH == 0x008u && SURVIVE == 0x00Cu) return life_from_triples(x, a, c, b);
    else if constexpr (BIRTH == 0x148u && SURVIVE == 0x034u) return morley_from_triples(x, a, c, b);
    else if constexpr (BIRTH == 0x048u && SURVIVE == 0x00Cu) return highlife_from_triples(x, a, c, b);
    else if constexpr (BIRTH == 0x1C8u && SURVIVE == 0x1D8u) return daynight_from_triples(x, a, c, b);
    else return next_static<BIRTH, SURVIVE>(x, add3(a, c, b));
}

// Rule known only at run time (any of the 2^18 B/S masks), branch free.  The five
// indicator planes e_u = [u == k] are rule independent; for each (x, t0) class the next
// state is OR_u (e_u & G[class][u]) with G = all-ones / all-zeros words expanded from the
// masks on the host (kept in the kernel's constant bank, one constant operand per LOP3).
// 18 of the 20 (class, u) pairs can occur.
struct RuleMasks {
    uint32_t d0[5];   // dead,  t0 = 0: birth bit 2u      (u = 0..4)
    uint32_t d1[4];   // dead,  t0 = 1: birth bit 2u + 1  (u = 0..3)
    uint32_t a0[4];   // alive, t0 = 0: survive bit 2u - 1 (u = 1..4), stored at [u - 1]
    uint32_t a1[5];   // alive, t0 = 1: survive bit 2u     (u = 0..4)
};

inline RuleMasks expand_rule(uint32_t birth, uint32_t survive) {
    RuleMasks m;
    for (int u = 0; u < 5; ++u) m.d0[u] = ((birth >> (2 * u)) & 1u) ? 0xFFFFFFFFu : 0u;
    for (int u = 0; u < 4; ++u) m.d1[u] = ((birth >> (2 * u + 1)) & 1u) ? 0xFFFFFFFFu : 0u;
    for (int u = 1; u < 5; ++u) m.a0[u - 1] = ((survive >> (2 * u - 1)) & 1u) ? 0xFFFFFFFFu : 0u;
    for (int u = 0; u < 5; ++u) m.a1[u] = ((survive >> (2 * u)) & 1u) ? 0xFFFFFFFFu : 0u;
    return m;
}

CA_HD uint32_t next_dynamic(uint32_t x, Sum9 s, const RuleMasks& r) {
    const uint32_t k0 = s.k0, t1 = s.t1, k1 = s.k1, t0 = s.t0;
    const uint32_t e0 = ~(k0 | t1 | k1);
    const uint32_t e1 = (k0 ^ t1) & ~k1;
    const uint32_t e2 = (k0 & t1 & ~k1) | (~k0 & ~t1 & k1);
    const uint32_t e3 = (k0 ^ t1) & k1;
    const uint32_t e4 = k0 & t1 & k1;
    const uint32_t gd0 = (e0 & r.d0[0]) | (e1 & r.d0[1]) | (e2 & r.d0[2]) | (e3 & r.d0[3]) |
                         (e4 & r.d0[4]);
    const uint32_t gd1 = (e0 & r.d1[0]) | (e1 & r.d1[1]) | (e2 & r.d1[2]) | (e3 & r.d1[3]);
    const uint32_t ga0 = (e1 & r.a0[0]) | (e2 & r.a0[1]) | (e3 & r.a0[2]) | (e4 & r.a0[3]);
    const uint32_t ga1 = (e0 & r.a1[0]) | (e1 & r.a1[1]) | (e2 & r.a1[2]) | (e3 & r.a1[3]) |
                         (e4 & r.a1[4]);
    const uint32_t dead = (t0 & gd1) | (~t0 & gd0);
    const uint32_t alive = (t0 & ga1) | (~t0 & ga0);
    return (x & alive) | (~x & dead);
}

// ---- weighted popcounts for the mcl.py SpeedDetector sums -------------------------------
// sum over set bits b of word v of b (0..31): 5 masked popcounts.
CA_HD uint32_t popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return (uint32_t)__builtin_popcount(v);
#endif
}
CA_HD uint32_t bit_index_sum(uint32_t v) {
    return popc32(v & 0xAAAAAAAAu) + 2u * popc32(v & 0xCCCCCCCCu) +
           4u * popc32(v & 0xF0F0F0F0u) + 8u * popc32(v & 0xFF00FF00u) +
           16u * popc32(v & 0xFFFF0000u);
}

// ---- carry-save column counts: N words of equal weight -> bit planes -------------------
// planes[p] bit b = bit p of (number of the N input words that have bit b set).  Built from
// 3:2 compressors (2 LOP3 each), ~2N instructions; lets one masked popcount per PLANE replace
// one per WORD (strip_lane_sums below).
CA_HD void full_add(uint32_t a, uint32_t b, uint32_t c, uint32_t& sum, uint32_t& carry) {
    sum = lop3<LUT_XOR3>(a, b, c);
    carry = lop3<LUT_MAJ>(a, b, c);
}

constexpr int csa_carries(int n) { return n <= 1 ? 0 : (n - 1) / 2 + ((n - 1) % 2); }
constexpr int csa_planes(int n) { return n <= 0 ? 0 : 1 + csa_planes(csa_carries(n)); }

template <int N>
struct CsaLevel {        // N same-weight words -> 1 word of that weight + csa_carries(N) carries
    static CA_HD void run(const uint32_t* a, uint32_t& plane, uint32_t* carry) {
        if constexpr (N == 1) {
            plane = a[0];
        } else if constexpr (N == 2) {
            plane = a[0] ^ a[1];
            carry[0] = a[0] & a[1];
        } else {
            constexpr int G = N / 3, L = N % 3;
            uint32_t nxt[G + L];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int g = 0; g < G; ++g) full_add(a[3 * g], a[3 * g + 1], a[3 * g + 2], nxt[g], carry[g]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int l = 0; l < L; ++l) nxt[G + l] = a[3 * G + l];
            CsaLevel<G + L>::run(nxt, plane, carry + G);
        }
    }
};

template <int N>
struct CsaTree {
    static constexpr int PLANES = csa_planes(N);
    static CA_HD void run(const uint32_t* a, uint32_t* planes) {
        if constexpr (N == 1) {
            planes[0] = a[0];
        } else {
            uint32_t carry[csa_carries(N)];
            CsaLevel<N>::run(a, planes[0], carry);
            CsaTree<csa_carries(N)>::run(carry, planes + 1);
        }
    }
};

// ---- window geometry of a centred square action window (carle/env.py:119-132) ------------
constexpr uint32_t window_col_mask_of(int w, int col0, int ah) {
    int lo = col0 - 32 * w; if (lo < 0) lo = 0;
    int hi = col0 + ah - 32 * w; if (hi > 32) hi = 32;
    if (hi <= lo) return 0u;
    return ((hi - lo == 32) ? 0xFFFFFFFFu : ((1u << (hi - lo)) - 1u)) << lo;
}

// ---- SpeedDetector partial sums (carle/mcl.py:773-779) of R rows x WPL words ----------------
// x[r][w] = row (row_base + r), columns [32w, 32w+32) of a 32*WPL-wide universe whose centred
// AWIN x AWIN action window starts at (ROW0, ROW0).  Adds to live (all cells), wl (cells inside
// the window), sh / sw (sum of row / column index over the cells OUTSIDE the window).
template <int WPL, int R, int AWIN>
CA_HD void strip_lane_sums(const uint32_t (&x)[R][WPL], int row_base, uint32_t& live,
                           uint32_t& sh, uint32_t& sw, uint32_t& wl) {
    constexpr int ROW0 = (32 * WPL - AWIN) / 2;
    uint32_t o[R * WPL];
    uint32_t inside = 0, outside = 0, wsum = 0, rsum = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < R; ++r) {
        const int row = row_base + r;
        const uint32_t rowmask = ((uint32_t)(row - ROW0) < (uint32_t)AWIN) ? 0xFFFFFFFFu : 0u;
        uint32_t rc = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int w = 0; w < WPL; ++w) {
            const uint32_t cm = window_col_mask_of(w, ROW0, AWIN);
            uint32_t out = x[r][w];
            if (cm != 0u) {
                const uint32_t ins = out & cm & rowmask;
                out ^= ins;
                inside += popc32(ins);
            }
            o[r * WPL + w] = out;
            const uint32_t pc = popc32(out);
            rc += pc;
            wsum += (uint32_t)w * pc;
        }
        outside += rc;
        rsum += (uint32_t)row * rc;
    }
    // sum over the words of bit_index_sum(word), one masked popcount per carry-save plane
    uint32_t planes[CsaTree<R * WPL>::PLANES];
    CsaTree<R * WPL>::run(o, planes);
    uint32_t bsum = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int p = 0; p < CsaTree<R * WPL>::PLANES; ++p) bsum += bit_index_sum(planes[p]) << p;
    live += outside + inside;
    wl += inside;
    sh += rsum;
    sw += 32u * wsum + bsum;
}

}  // namespace ca
