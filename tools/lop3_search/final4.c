// lop3_search/final4.c — the same question as final3.c with a FOUR-LOP3 final stage (y1, y2, y3, out).  The
// last two nodes are not enumerated: for every pair (p, q) of earlier signals the rule is decomposed as
// out = f(p, q, r), which fixes r up to a complement on every (p, q) class where the rule is not constant,
// and r (with the remaining rows as don't-cares) must be a LOP3 of three earlier signals.
// Morley B368/S245 and HighLife B36/S23: yes (ca::morley_from_triples, ca::highlife_from_triples: 8 LOP3
// per word instead of 11 / 10); Day & Night: not on the encoding the other rules use.
//   gcc -O3 -march=native -DBIRTH=0x148 -DSURV=0x034 -o final4 final4.c && ./final4 [encoding index 0..8]
// Output signal numbering: 0 = x, 1 = A, 2 = B, 3 = U, 4 = V, 5 = y1, 6 = y2; ea / eb / eu / ev = the
// encodings as 4-bit tables over L / H = 0..3; L<n>(a,b,c) = lop3 with immediate n on signals a, b, c.
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifndef BIRTH
#define BIRTH 0x148
#define SURV 0x034
#endif
static inline uint32_t lut3(uint32_t a, uint32_t b, uint32_t c, int lut) {
    uint32_t r = 0;
    for (int m = 0; m < 8; ++m) if (lut >> m & 1) {
        uint32_t t = 0xFFFFFFFFu;
        t &= (m & 4) ? a : ~a; t &= (m & 2) ? b : ~b; t &= (m & 1) ? c : ~c;
        r |= t;
    }
    return r;
}
static inline int is_fn(uint32_t target, uint32_t a, uint32_t b, uint32_t c, uint32_t care, int* lut) {
    int l = 0;
    for (int m = 0; m < 8; ++m) {
        uint32_t t = care;
        t &= (m & 4) ? a : ~a; t &= (m & 2) ? b : ~b; t &= (m & 1) ? c : ~c;
        if (!t) continue;
        uint32_t v = target & t;
        if (v != 0 && v != t) return 0;
        if (v) l |= 1 << m;
    }
    *lut = l;
    return 1;
}
static int inj(int a,int b){ int seen=0; for(int k=0;k<4;++k){int c=((a>>k)&1)|(((b>>k)&1)<<1); if(seen>>c&1) return 0; seen|=1<<c;} return ((a&1)==0)&&((b&1)==0); }

typedef struct { uint32_t tt; int a,b,c,lut; } Node;
static int gen_nodes(const uint32_t* sig, int n, Node* out) {
    // all distinct (up to complement) non-trivial LOP3 outputs over triples of sig[0..n)
    int cnt = 0;
    for (int a = 0; a < n; ++a) for (int b = a + 1; b < n; ++b) for (int c = b + 1; c < n; ++c)
        for (int l = 0; l < 128; ++l) {           // complement-canonical: LUT bit 7 = 0
            uint32_t t = lut3(sig[a], sig[b], sig[c], l);
            if (t == 0 || t == 0xFFFFFFFFu) continue;
            int dup = 0;
            for (int i = 0; i < n && !dup; ++i) if (t == sig[i] || t == ~sig[i]) dup = 1;
            for (int i = 0; i < cnt && !dup; ++i) if (t == out[i].tt || t == ~out[i].tt) dup = 1;
            if (dup) continue;
            out[cnt].tt = t; out[cnt].a = a; out[cnt].b = b; out[cnt].c = c; out[cnt].lut = l; ++cnt;
        }
    return cnt;
}
int main(int argc, char** argv) {
    int want_enc = argc > 1 ? atoi(argv[1]) : -1, enc = 0;
    static Node n1[4096], n2[8192];
    for (int ea = 0; ea < 16; ++ea) for (int eb = ea; eb < 16; ++eb) for (int eu = 0; eu < 16; ++eu) for (int ev = eu; ev < 16; ++ev) {
        if (!inj(ea, eb) || !inj(eu, ev)) continue;
        if (want_enc >= 0 && enc++ != want_enc) continue;
        uint32_t sig[8] = {0}, F = 0;
        for (int r = 0; r < 32; ++r) {
            int xb = r & 1, L = (r >> 1) & 3, H = (r >> 3) & 3, s9 = L + 2 * H;
            if (xb) sig[0] |= 1u << r;
            if (ea >> L & 1) sig[1] |= 1u << r;
            if (eb >> L & 1) sig[2] |= 1u << r;
            if (eu >> H & 1) sig[3] |= 1u << r;
            if (ev >> H & 1) sig[4] |= 1u << r;
            int on = xb ? (s9 >= 1 && ((SURV >> (s9 - 1)) & 1)) : ((BIRTH >> s9) & 1);
            if (on) F |= 1u << r;
        }
        int c1 = gen_nodes(sig, 5, n1);
        fprintf(stderr, "enc ea=%d eb=%d eu=%d ev=%d: %d y1 candidates\n", ea, eb, eu, ev, c1);
        long found = 0;
        for (int i1 = 0; i1 < c1 && found < 5; ++i1) {
            sig[5] = n1[i1].tt;
            int c2 = gen_nodes(sig, 6, n2);
            for (int i2 = 0; i2 < c2 && found < 5; ++i2) {
                sig[6] = n2[i2].tt;
                for (int p = 0; p < 7 && found < 5; ++p) for (int q = p + 1; q < 7 && found < 5; ++q) {
                    uint32_t cls[4], care = 0; int nc = 0; uint32_t ncl[4];
                    int bad = 0;
                    for (int m = 0; m < 4; ++m) {
                        uint32_t t = 0xFFFFFFFFu;
                        t &= (m & 2) ? sig[p] : ~sig[p]; t &= (m & 1) ? sig[q] : ~sig[q];
                        cls[m] = t;
                        uint32_t v = F & t;
                        if (t && v != 0 && v != t) { ncl[nc++] = t; care |= t; }
                    }
                    (void)bad; (void)cls;
                    if (nc == 0) continue;      // F is a function of (p,q) alone: would be found with fewer nodes
                    for (int pol = 0; pol < (1 << (nc - 1)); ++pol) {      // first class polarity fixed (r vs ~r)
                        uint32_t r = 0;
                        for (int k = 0; k < nc; ++k) {
                            uint32_t v = F & ncl[k];
                            if (k > 0 && (pol >> (k - 1) & 1)) v = ~F & ncl[k];
                            r |= v;
                        }
                        for (int a = 0; a < 7; ++a) for (int b = a + 1; b < 7; ++b) for (int c = b + 1; c < 7; ++c) {
                            int l3;
                            if (is_fn(r, sig[a], sig[b], sig[c], care, &l3)) {
                                uint32_t y3 = lut3(sig[a], sig[b], sig[c], l3);
                                int l4;
                                if (!is_fn(F, sig[p], sig[q], y3, 0xFFFFFFFFu, &l4)) continue;
                                printf("FOUND ea=%d eb=%d eu=%d ev=%d y1=L%d(%d,%d,%d) y2=L%d(%d,%d,%d) y3=L%d(%d,%d,%d) out=L%d(%d,%d,y3)\n",
                                       ea, eb, eu, ev, n1[i1].lut, n1[i1].a, n1[i1].b, n1[i1].c,
                                       n2[i2].lut, n2[i2].a, n2[i2].b, n2[i2].c, l3, a, b, c, l4, p, q);
                                fflush(stdout);
                                ++found;
                            }
                        }
                    }
                }
            }
        }
        fprintf(stderr, "enc done found=%ld\n", found);
    }
    return 0;
}
