/* smcmc_rng.h -- the random-draw stream shared by host and device.
 *
 * The reference draws from ROOT's global gRandom (TSimpleMCMC.H:172-179,
 * call sites :455,:691,:700,:713,:719).  An ensemble engine needs a
 * counter-based stream so that any (chain, step, slot) draw can be produced
 * independently on any GPU; this header defines that stream ONCE, in plain
 * C99-compatible code that compiles unchanged under gcc and nvcc, and whose
 * results are bit-identical on host and device:
 *
 *   - Philox4x32-10 (Salmon et al., SC'11) keyed by the 64-bit run seed,
 *     counter = (chain, step, slot, stream);
 *   - Rndm() in the open interval (0,1) like TRandom3 (never 0, never 1);
 *   - Gaus(0,1) by Box-Muller, TWO normals per Philox block: draw slots 2k and
 *     2k+1 are the cosine and the sine branch of the same 128 bits (one log,
 *     one sqrt, one argument reduction for both), taken from a sub-stream of
 *     their own (SMCMC_STREAM_NORMAL_PAIR) so that a uniform draw of slot 2k
 *     never shares bits with the normal of slot 2k+1; log / sin / cos are
 *     evaluated with
 *     polynomial kernels that use ONLY +,-,*,/ and sqrt, i.e. operations IEEE
 *     754 rounds identically on x86 and on sm_100a.  Every operation goes
 *     through the SMCMC_ADD/SUB/MUL/DIV macros, which on the device are the
 *     __dXXX_rn intrinsics (never contracted into FMAs) and on the host are
 *     plain operators (build with -ffp-contract=off; gcc on baseline x86-64
 *     has no FMA to contract into anyway).
 *
 * Slot map for one Metropolis step of an n-dimensional chain
 * (order of gRandom calls in TSimpleMCMC.H:709-724 then :455):
 *     slot i in [0,n) : proposal draw for dimension i
 *                       (Gaus(0,1), or Uniform(a,b) for a SetUniform() dim)
 *     slot n          : the Metropolis accept/reject Uniform()
 * stream = SMCMC_STREAM_STEP for these.
 */
#ifndef SMCMC_RNG_H_SEEN
#define SMCMC_RNG_H_SEEN

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define SMCMC_HD __host__ __device__ __forceinline__
#else
#define SMCMC_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define SMCMC_ADD(a, b) __dadd_rn((a), (b))
#define SMCMC_SUB(a, b) __dsub_rn((a), (b))
#define SMCMC_MUL(a, b) __dmul_rn((a), (b))
#define SMCMC_DIV(a, b) __ddiv_rn((a), (b))
#define SMCMC_SQRT(a) __dsqrt_rn((a))
#else
#define SMCMC_ADD(a, b) ((a) + (b))
#define SMCMC_SUB(a, b) ((a) - (b))
#define SMCMC_MUL(a, b) ((a) * (b))
#define SMCMC_DIV(a, b) ((a) / (b))
#define SMCMC_SQRT(a) sqrt((a))
#endif

enum {
    SMCMC_STREAM_STEP = 0,      /* proposal + accept draws of a MH step     */
    SMCMC_STREAM_HMC = 1,       /* momentum / epsilon / accept draws of HMC */
    SMCMC_STREAM_INPUT = 2      /* synthetic-input generators               */
};
/* OR-ed into the stream word of the Philox blocks that feed normal PAIRS: block
 * `pair` of that sub-stream yields the normals of draw slots 2*pair, 2*pair+1. */
#define SMCMC_STREAM_NORMAL_PAIR 0x100u

typedef struct smcmc_u32x4 {
    uint32_t v[4];
} smcmc_u32x4;

SMCMC_HD uint32_t smcmc_mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

/* Philox4x32-10.  ctr/key layout follows Random123 (philox.h) so its
 * published known-answer vectors apply (tests/test_rng.py). */
SMCMC_HD smcmc_u32x4 smcmc_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                         uint32_t c3, uint32_t k0,
                                         uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int round = 0; round < 10; ++round) {
        uint32_t hi0 = smcmc_mulhi32(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = smcmc_mulhi32(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    smcmc_u32x4 out;
    out.v[0] = c0; out.v[1] = c1; out.v[2] = c2; out.v[3] = c3;
    return out;
}

/* The 128 random bits of draw (chain, step, slot) of a stream. */
SMCMC_HD smcmc_u32x4 smcmc_draw_bits(uint64_t seed, uint32_t chain,
                                     uint32_t step, uint32_t slot,
                                     uint32_t stream) {
    return smcmc_philox4x32_10(chain, step, slot, stream,
                               (uint32_t)(seed & 0xffffffffu),
                               (uint32_t)(seed >> 32));
}

/* 52 random bits -> double in the open interval (0,1):  (k + 1/2) * 2^-52,
 * k < 2^52.  k + 1/2 has at most 53 significant bits, so the conversion, the
 * sum and the scaling are all exact: the smallest value is 2^-53, the largest
 * 1 - 2^-53, never 0 and never 1. */
SMCMC_HD double smcmc_bits_to_open01(uint32_t hi, uint32_t lo) {
    uint64_t k = (((uint64_t)hi << 32) | (uint64_t)lo) >> 12;
    return SMCMC_MUL(SMCMC_ADD((double)k, 0.5), 2.2204460492503130808e-16);
}

/* log(x) for finite x > 0, fdlibm-style reduction x = 2^e * m with
 * m in [sqrt(1/2), sqrt(2)), log(m) from the atanh series in s = f/(2+f).
 * Only exact integer manipulation and +,-,*,/ : bit-identical everywhere. */
SMCMC_HD double smcmc_det_log(double x) {
    const double ln2_hi = 6.93147180369123816490e-01;
    const double ln2_lo = 1.90821492927058770002e-10;
    const double Lg1 = 6.666666666666735130e-01;
    const double Lg2 = 3.999999999940941908e-01;
    const double Lg3 = 2.857142874366239149e-01;
    const double Lg4 = 2.222219843214978396e-01;
    const double Lg5 = 1.818357216161805012e-01;
    const double Lg6 = 1.531383769920937332e-01;
    const double Lg7 = 1.479819860511658591e-01;
    union { double d; uint64_t u; } cv;
    cv.d = x;
    int e = 0;
    if ((cv.u >> 52) == 0) {          /* subnormal: scale by 2^54 (exact) */
        cv.d = SMCMC_MUL(cv.d, 18014398509481984.0);
        e = -54;
    }
    e += (int)((cv.u >> 52) & 0x7ff) - 1023;
    uint64_t mant = cv.u & 0x000fffffffffffffull;
    /* m in [1,2); move the upper part to [sqrt(1/2),1) */
    if (mant >= 0x6a09e667f3bcdull) {
        cv.u = mant | 0x3fe0000000000000ull;   /* m/2 */
        e += 1;
    } else {
        cv.u = mant | 0x3ff0000000000000ull;
    }
    double f = SMCMC_SUB(cv.d, 1.0);
    double s = SMCMC_DIV(f, SMCMC_ADD(2.0, f));
    double z = SMCMC_MUL(s, s);
    double w = SMCMC_MUL(z, z);
    double t1 = SMCMC_MUL(w, SMCMC_ADD(Lg2, SMCMC_MUL(w, SMCMC_ADD(Lg4, SMCMC_MUL(w, Lg6)))));
    double t2 = SMCMC_MUL(z, SMCMC_ADD(Lg1, SMCMC_MUL(w, SMCMC_ADD(Lg3, SMCMC_MUL(w, SMCMC_ADD(Lg5, SMCMC_MUL(w, Lg7)))))));
    double R = SMCMC_ADD(t2, t1);
    double hfsq = SMCMC_MUL(0.5, SMCMC_MUL(f, f));
    double dk = (double)e;
    /* log(x) = k*ln2_hi - ((hfsq - (s*(hfsq+R) + k*ln2_lo)) - f) */
    double inner = SMCMC_ADD(SMCMC_MUL(s, SMCMC_ADD(hfsq, R)), SMCMC_MUL(dk, ln2_lo));
    return SMCMC_SUB(SMCMC_MUL(dk, ln2_hi),
                     SMCMC_SUB(SMCMC_SUB(hfsq, inner), f));
}

/* sin and cos of an angle |x| <= pi/4 (fdlibm kernel polynomials). */
SMCMC_HD double smcmc_det_ksin(double x) {
    const double S1 = -1.66666666666666324348e-01;
    const double S2 = 8.33333333332248946124e-03;
    const double S3 = -1.98412698298579493134e-04;
    const double S4 = 2.75573137070700676789e-06;
    const double S5 = -2.50507602534068634195e-08;
    const double S6 = 1.58969099521155010221e-10;
    double z = SMCMC_MUL(x, x);
    double v = SMCMC_MUL(z, x);
    double r = SMCMC_ADD(S2, SMCMC_MUL(z, SMCMC_ADD(S3, SMCMC_MUL(z, SMCMC_ADD(S4, SMCMC_MUL(z, SMCMC_ADD(S5, SMCMC_MUL(z, S6))))))));
    return SMCMC_ADD(x, SMCMC_MUL(v, SMCMC_ADD(S1, SMCMC_MUL(z, r))));
}

SMCMC_HD double smcmc_det_kcos(double x) {
    const double C1 = 4.16666666666666019037e-02;
    const double C2 = -1.38888888888741095749e-03;
    const double C3 = 2.48015872894767294178e-05;
    const double C4 = -2.75573143513906633035e-07;
    const double C5 = 2.08757232129817482790e-09;
    const double C6 = -1.13596475577881948265e-11;
    double z = SMCMC_MUL(x, x);
    double r = SMCMC_MUL(z, SMCMC_ADD(C1, SMCMC_MUL(z, SMCMC_ADD(C2, SMCMC_MUL(z, SMCMC_ADD(C3, SMCMC_MUL(z, SMCMC_ADD(C4, SMCMC_MUL(z, SMCMC_ADD(C5, SMCMC_MUL(z, C6)))))))))));
    double hz = SMCMC_MUL(0.5, z);
    double w = SMCMC_SUB(1.0, hz);
    /* 1 - (hz - z*r) with the fdlibm correction term for the rounding of w */
    return SMCMC_ADD(w, SMCMC_ADD(SMCMC_SUB(SMCMC_SUB(1.0, w), hz), SMCMC_MUL(z, r)));
}

/* cos(2*pi*u) and sin(2*pi*u) for u in (0,1) on the 2^-53 grid: exact octant
 * reduction.  `which`: 1 cosine only, 2 sine only, 3 both.
 * octant:  0     1     2      3      4      5     6     7     (a: reduced angle)
 * cos  :  cos a sin a -sin a -cos a -cos a -sin a sin a cos a
 * sin  :  sin a cos a  cos a  sin a -sin a -cos a -cos a -sin a */
SMCMC_HD void smcmc_det_sincos2pi(double u, int which, double* c, double* s) {
    const double quarter_pi = 7.85398163397448278999e-01;
    double t = SMCMC_MUL(u, 8.0);               /* exact */
    int oct = (int)t;                           /* 0..7 */
    double r = SMCMC_SUB(t, (double)oct);       /* exact, in [0,1) */
    if (oct & 1) r = SMCMC_SUB(1.0, r);         /* exact */
    double a = SMCMC_MUL(r, quarter_pi);        /* angle in [0, pi/4] */
    const int swap = ((oct + 1) >> 1) & 1;      /* octants 1, 2, 5, 6 */
    double ks = 0.0, kc = 0.0;
    if (which == 3 || (which == 1) == (swap != 0)) ks = smcmc_det_ksin(a);
    if (which == 3 || (which == 1) != (swap != 0)) kc = smcmc_det_kcos(a);
    if (which & 1) {
        double v = swap ? ks : kc;
        if (oct >= 2 && oct <= 5) v = -v;
        *c = v;
    }
    if (which & 2) {
        double v = swap ? kc : ks;
        if (oct >= 4) v = -v;
        *s = v;
    }
}

SMCMC_HD double smcmc_det_cos2pi(double u) {
    double c = 0.0, s = 0.0;
    smcmc_det_sincos2pi(u, 1, &c, &s);
    return c;
}

SMCMC_HD double smcmc_det_sin2pi(double u) {
    double c = 0.0, s = 0.0;
    smcmc_det_sincos2pi(u, 2, &c, &s);
    return s;
}

/* Rndm() of draw (chain, step, slot). */
SMCMC_HD double smcmc_uniform(uint64_t seed, uint32_t chain, uint32_t step,
                              uint32_t slot, uint32_t stream) {
    smcmc_u32x4 b = smcmc_draw_bits(seed, chain, step, slot, stream);
    return smcmc_bits_to_open01(b.v[0], b.v[1]);
}

/* The two halves of Box-Muller on one 128-bit block, separately callable (a kernel may
 * give them to two threads): the radius sqrt(-2 log u1) from words 0-1, the direction
 * (cos, sin)(2 pi u2) from words 2-3. */
SMCMC_HD double smcmc_normal_pair_radius(smcmc_u32x4 b) {
    double u1 = smcmc_bits_to_open01(b.v[0], b.v[1]);
    return SMCMC_SQRT(SMCMC_MUL(-2.0, smcmc_det_log(u1)));
}
SMCMC_HD void smcmc_normal_pair_direction(smcmc_u32x4 b, int which, double* c, double* s) {
    double u2 = smcmc_bits_to_open01(b.v[2], b.v[3]);
    smcmc_det_sincos2pi(u2, which, c, s);
}

/* Box-Muller on one 128-bit block: z0 = rad cos(2 pi u2), z1 = rad sin(2 pi u2),
 * rad = sqrt(-2 log u1).  `which` as smcmc_det_sincos2pi. */
SMCMC_HD void smcmc_normal_pair_from_bits(smcmc_u32x4 b, int which, double* z0, double* z1) {
    double rad = smcmc_normal_pair_radius(b);
    double c = 0.0, s = 0.0;
    smcmc_normal_pair_direction(b, which, &c, &s);
    if (which & 1) *z0 = SMCMC_MUL(rad, c);
    if (which & 2) *z1 = SMCMC_MUL(rad, s);
}

/* The Philox block behind the normals of draw slots 2*pair and 2*pair+1. */
SMCMC_HD smcmc_u32x4 smcmc_normal_pair_bits(uint64_t seed, uint32_t chain, uint32_t step,
                                            uint32_t pair, uint32_t stream) {
    return smcmc_draw_bits(seed, chain, step, pair, stream | SMCMC_STREAM_NORMAL_PAIR);
}

/* Gaus(0,1) of draw slots 2*pair (z0) and 2*pair+1 (z1) of (chain, step). */
SMCMC_HD void smcmc_normal_pair(uint64_t seed, uint32_t chain, uint32_t step,
                                uint32_t pair, uint32_t stream, double* z0, double* z1) {
    smcmc_normal_pair_from_bits(smcmc_normal_pair_bits(seed, chain, step, pair, stream), 3, z0, z1);
}

/* Gaus(0,1) of draw (chain, step, slot): one branch of its pair -- the cosine branch for an
 * even slot, the sine branch for an odd one.  Same operations as the pair routine; written out
 * so that a thread evaluates exactly one of the two polynomial kernels. */
SMCMC_HD double smcmc_normal(uint64_t seed, uint32_t chain, uint32_t step,
                             uint32_t slot, uint32_t stream) {
    const double quarter_pi = 7.85398163397448278999e-01;
    smcmc_u32x4 b = smcmc_normal_pair_bits(seed, chain, step, slot >> 1, stream);
    double rad = smcmc_normal_pair_radius(b);
    double u2 = smcmc_bits_to_open01(b.v[2], b.v[3]);
    double t = SMCMC_MUL(u2, 8.0);
    int oct = (int)t;
    double r = SMCMC_SUB(t, (double)oct);
    if (oct & 1) r = SMCMC_SUB(1.0, r);
    double a = SMCMC_MUL(r, quarter_pi);
    const int odd = (int)(slot & 1u);
    const int swap = ((oct + 1) >> 1) & 1;
    double v = (swap ^ odd) ? smcmc_det_ksin(a) : smcmc_det_kcos(a);
    if (odd ? (oct >= 4) : (oct >= 2 && oct <= 5)) v = -v;
    return SMCMC_MUL(rad, v);
}

#endif
