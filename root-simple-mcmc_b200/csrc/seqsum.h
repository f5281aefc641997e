/* seqsum.h -- exact emulation of "add the same weight n times".
 *
 * The reference fills its simulated histograms one event at a time:
 * TH1::Fill(x, w) does content[bin] += w in event order
 * (example/FakeLikelihood.H:198-215).  All events of one weight class carry
 * the SAME weight for a given parameter point (SystematicCorrection.H:81-117
 * depends on the event only through Type and MuDk), so a bin's content is
 *        fl(...fl(fl(0 + w) + w)... + w)      (n additions)
 * for each class in turn.  The device counts events per (class, bin) with
 * exact integers and this routine reproduces the reference's rounded running
 * sum bit for bit from the count, in O(log n) instead of O(n) additions.
 *
 * Why jumping is exact: while the running sum s stays inside one binade
 * [2^e, 2^(e+1)) its ulp u is constant and s is a multiple of u, so
 * fl(s + w) - s is the same multiple of u at every step (round-to-nearest;
 * in the tie case w = q*u + u/2 ties-to-even makes s even after one step and
 * the increment constant from then on).  So once two consecutive increments
 * agree, the next m sums are s + m*inc exactly, as long as they stay below
 * the top of the binade.  Near the top the routine falls back to real
 * additions, which also handle the change of ulp.
 *
 * Plain C, host and device; tests/test_seqsum.py checks it against the naive
 * loop.
 */
#ifndef SMCMC_SEQSUM_H_SEEN
#define SMCMC_SEQSUM_H_SEEN

#include <stdint.h>
#include "smcmc_rng.h"   /* SMCMC_HD, SMCMC_ADD/SUB/MUL/DIV */

SMCMC_HD uint32_t smcmc_exponent_bits(double x) {
    union { double d; uint64_t u; } cv;
    cv.d = x;
    return (uint32_t)((cv.u >> 52) & 0x7ff);
}

SMCMC_HD double smcmc_seq_add(double s, double w, uint32_t n) {
    if (!(w > 0.0) || !(s >= 0.0)) {
        /* not the histogram case (negative, zero or NaN weight): plain loop */
        for (uint32_t i = 0; i < n; ++i) s = SMCMC_ADD(s, w);
        return s;
    }
    while (n > 0) {
        double s1 = SMCMC_ADD(s, w);
        if (s1 == s) return s;            /* w is below half an ulp: stuck */
        if (n < 4) { s = s1; --n; continue; }
        double s2 = SMCMC_ADD(s1, w);
        uint32_t e0 = smcmc_exponent_bits(s);
        if (e0 != 0 && e0 < 0x7fe && e0 == smcmc_exponent_bits(s2)) {
            double d1 = SMCMC_SUB(s1, s);     /* exact: same binade */
            double d2 = SMCMC_SUB(s2, s1);
            if (d1 == d2) {
                union { double d; uint64_t u; } top;
                top.u = (uint64_t)(e0 + 1) << 52;           /* 2^(e+1) */
                double room = SMCMC_SUB(top.d, s2);         /* exact */
                double mf = SMCMC_DIV(room, d1);
                if (mf > 4.0) {
                    mf = SMCMC_SUB(mf, 2.0);
                    uint32_t left = n - 2;
                    uint32_t m = (mf >= 4294967040.0) ? left : (uint32_t)mf;
                    if (m > left) m = left;
                    s = SMCMC_ADD(s2, SMCMC_MUL((double)m, d1));   /* exact */
                    n -= 2 + m;
                    continue;
                }
                /* within a few steps of the top of the binade: plain additions (the
                 * reference's own operation) up to the crossing, without the set-up above */
                s = s2;
                n -= 2;
                while (n > 0 && smcmc_exponent_bits(s) == e0) {
                    double t = SMCMC_ADD(s, w);
                    if (t == s) return s;
                    s = t;
                    --n;
                }
                continue;
            }
        }
        s = s2;
        n -= 2;
    }
    return s;
}

#endif
