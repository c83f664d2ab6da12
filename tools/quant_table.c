// Feasibility check for a table-driven waterfall quantiser ("Sketch for (ii)"). Not part of the product.
//
// The reference quantiser (src/fft_impl.cpp:14-44) is, per power offset, a function Q of the 31 non-sign bits of
// |X|^2. This tool builds, for one offset, a 2048-entry table indexed by bits >> 20 (exponent + 3 mantissa bits):
//     base = Q(first value of the cell), [lo, hi) = the band around the cell's single step inside which Q may
//     wiggle (rounding makes the polynomial non-monotone by an ulp) and the exact arithmetic has to be used,
// and then checks EXHAUSTIVELY (all 2^31 inputs) that
//     Q(b) == base                 for b <  lo
//     Q(b) == (base + 1) & 0xFF    for b >= hi
// Usage: gcc -O2 -ffp-contract=off -fopenmp tools/quant_table.c -o /tmp/quant_table && /tmp/quant_table [offset_lo offset_hi]
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

// op-for-op the reference arithmetic (every float op separately rounded: build with -ffp-contract=off)
static inline int Q(uint32_t bits, int off) {
    float log_val = (float)((int)((bits >> 23) & 0xFF) - 128) + (float)off;
    uint32_t mb = (bits & ~(255u << 23)) + (127u << 23);
    float m = u2f(mb);
    float poly = ((-0.34484843f * m) + 2.02466578f) * m - 0.67487759f;
    float v = ((log_val + poly) * 0.3010299956639812f) * 20.f + 127.f;
    v = v > -128.f ? v : -128.f;
    return (int)v & 0xFF;
}

typedef struct { uint32_t lo, hi; int base; } Cell;

static void build(Cell *tab, int off, int scan) {
    for (uint32_t c = 0; c < 2048; c++) {
        const uint32_t b0 = c << 20, b1 = b0 + (1u << 20);
        const int q0 = Q(b0, off);
        // bisection for "a" first index where Q != q0 (exact if Q were monotone), then widen by `scan` ulps
        uint32_t a = b0, b = b1;  // invariant: Q(a) == q0; b = b1 or Q(b) != q0
        if (Q(b1 - 1, off) == q0) { a = b1 - 1; b = b1; }
        while (b - a > 1) {
            uint32_t mid = a + (b - a) / 2;
            if (Q(mid, off) == q0) a = mid; else b = mid;
        }
        uint32_t lo = b, hi = b;
        uint32_t s0 = b > b0 + scan ? b - scan : b0, s1 = b + scan < b1 ? b + scan : b1;
        for (uint32_t x = s0; x < s1; x++) {
            if (Q(x, off) != q0 && x < lo) lo = x;      // first deviation
            if (Q(x, off) == q0 && x + 1 > hi) hi = x + 1;  // one past the last base value
        }
        if (lo > hi) lo = hi;
        tab[c].lo = lo; tab[c].hi = hi; tab[c].base = q0;
    }
}

int main(int argc, char **argv) {
    int o0 = argc > 2 ? atoi(argv[1]) : 20, o1 = argc > 2 ? atoi(argv[2]) : 20;
    for (int off = o0; off <= o1; off++) {
        static Cell tab[2048];
        build(tab, off, 64);
        unsigned long long bad = 0, band = 0;
        uint32_t widest = 0;
        for (int c = 0; c < 2048; c++) {
            band += tab[c].hi - tab[c].lo;
            if (tab[c].hi - tab[c].lo > widest) widest = tab[c].hi - tab[c].lo;
        }
#pragma omp parallel for reduction(+ : bad) schedule(dynamic, 16)
        for (int c = 0; c < 2048; c++) {
            const uint32_t b0 = (uint32_t)c << 20;
            const Cell t = tab[c];
            for (uint32_t i = 0; i < (1u << 20); i++) {
                const uint32_t b = b0 + i;
                if (b >= t.lo && b < t.hi) continue;  // exact path on the device
                const int want = Q(b, off);
                const int got = (t.base + (b >= t.lo)) & 0xFF;
                bad += (want != got);
            }
        }
        printf("offset %2d: mismatches outside the bands %llu of 2^31; inputs inside a band %llu (widest band %u ulps)\n", off, bad,
               band, widest);
        fflush(stdout);
    }
    return 0;
}
