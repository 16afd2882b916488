// lop3_search/final3.c — does a Life-like rule F(x, L, H) (x = the cell, L / H = the column sums of the
// low / high planes of the three row triples, sum9 = L + 2H; ca_core.cuh) have a THREE-LOP3 network on
// (x, A, B, U, V), where (A, B) / (U, V) is any injective 2-bit encoding of L / H (one LOP3 each from the
// triples)?  Exhaustive: y1 = any LUT of any 3 inputs, y2 = any LUT of any 3 of the 6 signals, out = any
// function of any 3 of the 7 signals.  Conway's Life: yes, on A = [L in {1,2}], B = L & 1 (same for H) --
// ca::life_from_triples, 7 LOP3 per word instead of 8; not on the binary encoding; Morley, HighLife,
// Day & Night: no.  (A two-LOP3 final stage exists for none of them.)
//   gcc -O3 -march=native -DBIRTH=0x008 -DSURV=0x00C -o final3 final3.c && ./final3
// BIRTH bit n: a dead cell with sum9 = n is born; SURV bit n: a live cell with n live neighbours survives.
// brute force: can Life's rule F(x,t0,k0,u,v) be computed with 3 LOP3 for some encoding (u,v) of H?
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#ifndef BIRTH
#define BIRTH 0x148
#define SURV 0x034
#endif
// truth tables over 5 inputs = 32 rows -> uint32
static uint32_t in_tt[5];
static inline uint32_t lut3(uint32_t a, uint32_t b, uint32_t c, int lut) {
    uint32_t r = 0;
    for (int m = 0; m < 8; ++m) if (lut >> m & 1) {
        uint32_t t = 0xFFFFFFFFu;
        t &= (m & 4) ? a : ~a; t &= (m & 2) ? b : ~b; t &= (m & 1) ? c : ~c;
        r |= t;
    }
    return r;
}
// is target a function of (a,b,c)?  rows with same (a,b,c) must agree
static inline int is_fn(uint32_t target, uint32_t a, uint32_t b, uint32_t c, uint32_t care) {
    for (int m = 0; m < 8; ++m) {
        uint32_t t = care;
        t &= (m & 4) ? a : ~a; t &= (m & 2) ? b : ~b; t &= (m & 1) ? c : ~c;
        if (t && (target & t) != 0 && (target & t) != t) return 0;
    }
    return 1;
}
static int inj(int a,int b){ int seen=0; for(int k=0;k<4;++k){int c=((a>>k)&1)|(((b>>k)&1)<<1); if(seen>>c&1) return 0; seen|=1<<c;} /* canonical: value at k=0 is 00 */ return ((a&1)==0)&&((b&1)==0); }
int main() {
    // rows indexed by bits: x=bit0, L=bits1-2 (0..3), H=bits3-4 (0..3)
    // encodings: (t0,k0) fixed as L bits; (u,v) = any two functions of H (16 x 16)
    int found = 0;
    for (int ea = 0; ea < 16; ++ea) for (int eb = ea; eb < 16; ++eb) for (int eu = 0; eu < 16; ++eu) for (int ev = eu; ev < 16; ++ev) { if (!inj(ea,eb) || !inj(eu,ev)) continue;
        uint32_t x = 0, t0 = 0, k0 = 0, u = 0, v = 0, target = 0;
        for (int r = 0; r < 32; ++r) {
            int xb = r & 1, L = (r >> 1) & 3, H = (r >> 3) & 3;
            int s9 = L + 2 * H;
            if (xb) x |= 1u << r;
            if (ea >> L & 1) t0 |= 1u << r;
            if (eb >> L & 1) k0 |= 1u << r;
            if (eu >> H & 1) u |= 1u << r;
            if (ev >> H & 1) v |= 1u << r;
            { int on = xb ? (s9 >= 1 && ((SURV >> (s9 - 1)) & 1)) : ((BIRTH >> s9) & 1); if (on) target |= 1u << r; }
        }
        // (u,v) must at least distinguish H=0,1,2 where needed; just search
        uint32_t sig[7] = {x, t0, k0, u, v, 0, 0};
        for (int a = 0; a < 5; ++a) for (int b = a + 1; b < 5; ++b) for (int c = b + 1; c < 5; ++c)
        for (int l1 = 0; l1 < 256; ++l1) {
            sig[5] = lut3(sig[a], sig[b], sig[c], l1);
            for (int d = 0; d < 6; ++d) for (int e = d + 1; e < 6; ++e) for (int f = e + 1; f < 6; ++f)
            for (int l2 = 0; l2 < 256; ++l2) {
                sig[6] = lut3(sig[d], sig[e], sig[f], l2);
                for (int g = 0; g < 7; ++g) for (int h = g + 1; h < 7; ++h) for (int i = h + 1; i < 7; ++i) {
                    if (i < 5) continue;   // must use at least one intermediate
                    if (is_fn(target, sig[g], sig[h], sig[i], 0xFFFFFFFFu)) {
                        printf("FOUND ea=%d eb=%d eu=%d ev=%d y1=L%d(%d,%d,%d) y2=L%d(%d,%d,%d) out(%d,%d,%d)\n", ea, eb, eu, ev, l1, a, b, c, l2, d, e, f, g, h, i);
                        found++;
                        if (found > 20) return 0;
                    }
                }
            }
        }
        fprintf(stderr, "enc %d %d %d %d done\n", ea, eb, eu, ev);
    }
    printf("found=%d\n", found);
    return 0;
}
