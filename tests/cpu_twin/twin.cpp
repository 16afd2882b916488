// Host build of carle_b200/csrc/ca_core.cuh for the CPU test-suite ONLY: lets the tests
// pin the bit-sliced arithmetic (row triples, 3x3 carry-save sum, static and run-time
// rule tables, seam handling) against the oracle without a GPU.  Not part of the product.
#include <cstdint>
#include <vector>
#include "../../carle_b200/csrc/ca_core.cuh"

namespace {
constexpr uint32_t kLifeB = 0x008, kLifeS = 0x00C, kMorleyB = 0x148, kMorleyS = 0x034,
                   kHighB = 0x048, kHighS = 0x00C, kDayB = 0x1C8, kDayS = 0x1D8;

// the path the kernels take: rule from the three row triples (Life: the 7-LOP3 network)
uint32_t apply_rule_triples(int mode, uint32_t x, ca::Triple a, ca::Triple c, ca::Triple b,
                            const ca::RuleMasks& m) {
    switch (mode) {
        case 1: return ca::next_static_triples<kLifeB, kLifeS>(x, a, c, b);
        case 2: return ca::next_static_triples<kMorleyB, kMorleyS>(x, a, c, b);
        case 3: return ca::next_static_triples<kHighB, kHighS>(x, a, c, b);
        case 4: return ca::next_static_triples<kDayB, kDayS>(x, a, c, b);
        default: return ca::next_dynamic(x, ca::add3(a, c, b), m);
    }
}

uint32_t apply_rule(int mode, uint32_t x, ca::Sum9 s, const ca::RuleMasks& m) {
    switch (mode) {
        case 1: return ca::next_static<kLifeB, kLifeS>(x, s);
        case 2: return ca::next_static<kMorleyB, kMorleyS>(x, s);
        case 3: return ca::next_static<kHighB, kHighS>(x, s);
        case 4: return ca::next_static<kDayB, kDayS>(x, s);
        default: return ca::next_dynamic(x, s, m);
    }
}
}  // namespace

extern "C" {

// bit p of the result = next state for (t0,k0,t1,k1,x) = bits 0..4 of p
uint32_t twin_rule_word(uint32_t birth, uint32_t survive, int mode) {
    ca::Sum9 s{0xAAAAAAAAu, 0xCCCCCCCCu, 0xF0F0F0F0u, 0xFF00FF00u};
    return apply_rule(mode, 0xFFFF0000u, s, ca::expand_rule(birth, survive));
}

// A built-in rule from the row triples (the path the kernels take) on all 2^7 inputs (x, lo/hi of
// the three triples): bit p of the result = next state for input
// p = x | lo_a << 1 | lo_c << 2 | lo_b << 3 | hi_a << 4 | hi_c << 5 | hi_b << 6  (four words of 32 inputs)
void twin_rule_triples(int mode, uint32_t out[4]) {
    const ca::RuleMasks none = ca::expand_rule(0, 0);
    for (int q = 0; q < 4; ++q) {
        uint32_t v[7] = {0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 32; ++i) {
            const int p = 32 * q + i;
            for (int k = 0; k < 7; ++k) if ((p >> k) & 1) v[k] |= 1u << i;
        }
        const ca::Triple a{v[1], v[4]}, c{v[2], v[5]}, b{v[3], v[6]};
        out[q] = apply_rule_triples(mode, v[0], a, c, b, none);
    }
}

// number of (rule, class) disagreements of next_dynamic with the definition, all 2^18 rules
long twin_exhaustive_dynamic() {
    long bad = 0;
    for (uint32_t b = 1; b < 512; ++b)
        for (uint32_t sv = 1; sv < 512; ++sv) {
            uint32_t got = twin_rule_word(b, sv, 0);
            for (int p = 0; p < 32; ++p) {
                int t0 = p & 1, k0 = (p >> 1) & 1, t1 = (p >> 2) & 1, k1 = (p >> 3) & 1, x = p >> 4;
                int sum9 = t0 + 2 * (k0 + t1) + 4 * k1;
                if ((x == 0 && sum9 > 8) || (x == 1 && sum9 < 1)) continue;   // unreachable
                int want = x ? ((sv >> (sum9 - 1)) & 1) : ((b >> sum9) & 1);
                if ((int)((got >> p) & 1) != want) ++bad;
            }
        }
    return bad;
}

// one generation on packed grids [n][h][wpr]; same word-level logic as step_generic_kernel
void twin_step(const uint32_t* in, uint32_t* out, long n, int h, int w, uint32_t birth,
               uint32_t survive, int mode) {
    const int wpr = (w + 31) / 32, tail = w & 31;
    const uint32_t tailmask = tail ? ((1u << tail) - 1u) : 0xFFFFFFFFu;
    const ca::RuleMasks masks = ca::expand_rule(birth, survive);
    for (long inst = 0; inst < n; ++inst) {
        const uint32_t* base = in + inst * (long)h * wpr;
        for (int r = 0; r < h; ++r)
            for (int c = 0; c < wpr; ++c) {
                const int wl = c == 0 ? wpr - 1 : c - 1, wr = c == wpr - 1 ? 0 : c + 1;
                ca::Triple t[3];
                uint32_t centre = 0;
                for (int d = 0; d < 3; ++d) {
                    int rr = r + d - 1;
                    rr = rr < 0 ? h - 1 : (rr >= h ? 0 : rr);
                    const uint32_t* row = base + (long)rr * wpr;
                    uint32_t xl = row[wl], xc = row[c], xr = row[wr];
                    uint32_t prev = (c == 0 && tail) ? (xl << (32 - tail)) : xl;
                    uint32_t west = ca::west(prev, xc);
                    uint32_t east = (c == wpr - 1 && tail)
                                        ? ((xc >> 1) | ((xr & 1u) << (tail - 1)))
                                        : ca::east(xc, xr);
                    t[d] = ca::row_triple(west, xc, east);
                    if (d == 1) centre = xc;
                }
                uint32_t nx = apply_rule_triples(mode, centre, t[0], t[1], t[2], masks);
                if (c == wpr - 1) nx &= tailmask;
                out[(inst * (long)h + r) * wpr + c] = nx;
            }
    }
}

uint32_t twin_bit_index_sum(uint32_t v) { return ca::bit_index_sum(v); }

}  // extern "C"

// column counts of n random-ish words through the carry-save tree vs a plain count
template <int N>
static int csa_bad(const uint32_t* words) {
    uint32_t planes[ca::CsaTree<N>::PLANES];
    ca::CsaTree<N>::run(words, planes);
    int bad = 0;
    for (int b = 0; b < 32; ++b) {
        int want = 0, got = 0;
        for (int i = 0; i < N; ++i) want += (words[i] >> b) & 1u;
        for (int p = 0; p < ca::CsaTree<N>::PLANES; ++p) got += ((planes[p] >> b) & 1u) << p;
        bad += want != got;
    }
    return bad;
}

extern "C" int twin_csa_bad(const uint32_t* words, int n) {
    switch (n) {
        case 1: return csa_bad<1>(words);   case 2: return csa_bad<2>(words);
        case 3: return csa_bad<3>(words);   case 4: return csa_bad<4>(words);
        case 5: return csa_bad<5>(words);   case 7: return csa_bad<7>(words);
        case 8: return csa_bad<8>(words);   case 12: return csa_bad<12>(words);
        case 16: return csa_bad<16>(words); case 32: return csa_bad<32>(words);
        case 64: return csa_bad<64>(words);
    }
    return -1;
}

// SpeedDetector sums of one packed instance [32*WPL][WPL], accumulated strip by strip and
// lane by lane exactly as step_strip_kernel does (lane L of strip q holds rows q*32R + L*R ..)
template <int WPL, int R, int AWIN>
static void strip_sums_of(const uint32_t* inst, uint32_t out[4]) {
    uint32_t live = 0, sh = 0, sw = 0, wl = 0;
    for (int q = 0; q < WPL / R; ++q)
        for (int lane = 0; lane < 32; ++lane) {
            uint32_t x[R][WPL];
            const int row_base = q * 32 * R + lane * R;
            for (int r = 0; r < R; ++r)
                for (int w = 0; w < WPL; ++w) x[r][w] = inst[(row_base + r) * WPL + w];
            ca::strip_lane_sums<WPL, R, AWIN>(x, row_base, live, sh, sw, wl);
        }
    out[0] = live; out[1] = sh; out[2] = sw; out[3] = wl;
}

extern "C" int twin_strip_sums(const uint32_t* inst, int wpl, int r, int awin, uint32_t out[4]) {
    if (wpl == 8 && r == 2 && awin == 64) strip_sums_of<8, 2, 64>(inst, out);
    else if (wpl == 8 && r == 4 && awin == 64) strip_sums_of<8, 4, 64>(inst, out);
    else if (wpl == 8 && r == 8 && awin == 64) strip_sums_of<8, 8, 64>(inst, out);
    else if (wpl == 4 && r == 2 && awin == 32) strip_sums_of<4, 2, 32>(inst, out);
    else if (wpl == 4 && r == 4 && awin == 32) strip_sums_of<4, 4, 32>(inst, out);
    else if (wpl == 2 && r == 2 && awin == 32) strip_sums_of<2, 2, 32>(inst, out);
    else if (wpl == 4 && r == 1 && awin == 96) strip_sums_of<4, 1, 96>(inst, out);
    else return -1;
    return 0;
}

