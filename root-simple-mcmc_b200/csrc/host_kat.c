/* host_kat.c -- host-side entry points onto the header-only pieces shared by
 * host and device (smcmc_rng.h, seqsum.h), so that the CPU test-suite can
 * check them without a GPU (known-answer tests for Philox, host/device
 * identical normal transform, exact sequential-sum emulation).  Built as
 * libsmcmc_hostkat.so by csrc/Makefile. */
#include <stddef.h>
#include <stdint.h>
#include "smcmc_rng.h"
#include "seqsum.h"
#include "hmc_order.h"

void smcmc_kat_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
    smcmc_u32x4 r = smcmc_philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    for (int i = 0; i < 4; ++i) out[i] = r.v[i];
}
double smcmc_kat_uniform(uint64_t seed, uint32_t chain, uint32_t step, uint32_t slot, uint32_t stream) {
    return smcmc_uniform(seed, chain, step, slot, stream);
}
double smcmc_kat_normal(uint64_t seed, uint32_t chain, uint32_t step, uint32_t slot, uint32_t stream) {
    return smcmc_normal(seed, chain, step, slot, stream);
}
void smcmc_kat_normals(uint64_t seed, uint32_t chain, uint32_t step0, uint32_t nsteps,
                       uint32_t nslots, double* out) {
    for (uint32_t s = 0; s < nsteps; ++s)
        for (uint32_t k = 0; k < nslots; ++k)
            out[(size_t)s * nslots + k] = smcmc_normal(seed, chain, step0 + s, k, SMCMC_STREAM_STEP);
}
double smcmc_kat_det_log(double x) { return smcmc_det_log(x); }
double smcmc_kat_det_cos2pi(double u) { return smcmc_det_cos2pi(u); }
double smcmc_kat_det_sin2pi(double u) { return smcmc_det_sin2pi(u); }
void smcmc_kat_normal_pair(uint64_t seed, uint32_t chain, uint32_t step, uint32_t pair, uint32_t stream, double* out2) {
    smcmc_normal_pair(seed, chain, step, pair, stream, out2, out2 + 1);
}
double smcmc_kat_bits_to_open01(uint32_t hi, uint32_t lo) { return smcmc_bits_to_open01(hi, lo); }
double smcmc_kat_seq_add(double s, double w, uint32_t n) { return smcmc_seq_add(s, w, n); }
double smcmc_kat_seq_add_naive(double s, double w, uint32_t n) {
    for (uint32_t i = 0; i < n; ++i) s = s + w;
    return s;
}

/* the chains of an HMC ensemble in order of trajectory length (hmc_order.h) */
long long smcmc_kat_hmc_order(const int* steps, int chains, int maxSteps, int tile, int* order, int* tiles, int* scratch) {
    return smcmc_hmc_order(steps, chains, maxSteps, tile, order, tiles, scratch);
}
