/* oracle/chain_api.h -- the single-chain C interface that BOTH CPU checkers
 * export: oracle/_ref/libsmcmc_ref.so (the reference's own headers compiled
 * unmodified against oracle/rootshim, prefix ref_) and
 * oracle/_build/libsmcmc_oracle.so (the stand-alone restatement in
 * oracle/smcmc_oracle.cc, prefix orc_).  TEST INFRASTRUCTURE ONLY: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference leg may load these libraries.  The product library
 * (libsmcmc_b200.so) never does.
 *
 * One handle = one chain = one sMCMC::TSimpleMCMC<L, TProposeAdaptiveStep>
 * (TSimpleMCMC.H:185-590, :640-1977) driven by the injected counter-based
 * stream of include/smcmc_rng.h: Step number s of chain c consumes slots
 * 0..n of (seed, c, s) in the order the reference calls gRandom.
 */
#ifndef SMCMC_ORACLE_CHAIN_API_H
#define SMCMC_ORACLE_CHAIN_API_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Likelihood kinds. */
enum {
    ORC_LLH_UNIT_GAUSS = 0, /* -1/2 sum x^2, the doc example TSimpleMCMC.H:111-120 */
    ORC_LLH_DUMMY = 1,      /* TDummyLogLikelihood.H:21-31, dense error matrix  */
    ORC_LLH_HORRIFIC = 2,   /* THorrificLogLikelihood.H:26-38                   */
    ORC_LLH_ASYM = 3,       /* TAsymLogLikelihood.H:20-31                       */
    ORC_LLH_FAKE = 4,       /* example/FakeLikelihood.H:47-81                   */
    ORC_LLH_HARD = 6,       /* THardLogLikelihood.H:57-91 (Rosenbrock, with gradient) */
    ORC_LLH_FAKE2 = 7,      /* example2/FakeLikelihood.H:58-118, 222-289 (events and data
                               set with *_chain_set_fake; the exposure is not used)   */
    ORC_LLH_CONSTRAINED = 8, /* example4/TConstrainedLikelihood.H:26-46 with the priors of its
                               Init() (:55-110): 25 dimensions, a constraint on the sum        */
    ORC_LLH_UNBINNED = 5    /* NOT in the reference: the unbinned
                               mixture likelihood of BASELINE.json configs[4], defined in
                               include/smcmc_b200.h; events set with *_chain_set_fake */
};

/* Scalar proposal settings: the TProposeAdaptiveStep setter each one calls. */
enum {
    ORC_SET_SIGMA = 0,                /* SetSigma                :775  */
    ORC_SET_TARGET_ACCEPTANCE = 1,    /* SetTargetAcceptance     :977  */
    ORC_SET_ACCEPTANCE_WINDOW = 2,    /* SetAcceptanceWindow     :982  */
    ORC_SET_ACCEPTANCE_RIGIDITY = 3,  /* SetAcceptanceRigidity   :1002 */
    ORC_SET_ACCEPTANCE_DEWEIGHT = 4,  /* SetAcceptanceUpdateDeweighting :987 */
    ORC_SET_COVARIANCE_WINDOW = 5,    /* SetCovarianceWindow     :914  */
    ORC_SET_COVARIANCE_DEWEIGHT = 6,  /* SetCovarianceUpdateDeweighting :927 */
    ORC_SET_COVARIANCE_FROZEN = 7,    /* SetCovarianceFrozen     :937  */
    ORC_SET_COVARIANCE_TRIALS = 8,    /* SetCovarianceTrials     :947  */
    ORC_SET_CENTER_TRIALS = 9,        /* SetEstimatedCenterTrials :747 */
    ORC_SET_NEXT_UPDATE = 10,         /* SetNextUpdate           :992  */
    ORC_SET_MAX_CORRELATION = 11,     /* SetMaximumCorrelation   :909  */
    ORC_SET_STEP_RMS_WINDOW = 12      /* TSimpleMCMC::SetStepRMSWindow :511 */
};

/* Layout of the scalar block returned by *_chain_get_state. */
enum {
    ORC_ST_SIGMA = 0, ORC_ST_ACCEPTANCE, ORC_ST_ACCEPTANCE_TRIALS,
    ORC_ST_ACCEPTANCE_WINDOW, ORC_ST_ACCEPTANCE_RIGIDITY,
    ORC_ST_TARGET_ACCEPTANCE, ORC_ST_TRIALS, ORC_ST_SUCCESSES,
    ORC_ST_NEXT_UPDATE, ORC_ST_COVARIANCE_TRIALS, ORC_ST_COVARIANCE_WINDOW,
    ORC_ST_CENTER_TRIALS, ORC_ST_COVARIANCE_TRACE, ORC_ST_SIGMA_TRACE,
    ORC_ST_STEP_RMS, ORC_ST_ACCEPTED_LLH, ORC_ST_PROPOSED_LLH,
    ORC_ST_TOTAL_STEPS, ORC_ST_LLH_CALLS, ORC_ST_COUNT
};

/* The reference's MC event record, example/Simulated.H:7-14 (48 bytes). */
typedef struct orc_event {
    double Mass;
    int32_t Type;
    int32_t pad0_;
    double Separation;
    int32_t MuDk;
    int32_t pad1_;
    double TrueMass;
    double TrueMassSigma;
} orc_event;

#define ORC_DECLARE(P)                                                        \
    void* P##chain_create(int kind, int dim, uint64_t seed, uint32_t chain);  \
    /* TSimpleMCMC<L, TProposeVAATStep> (TProposeVAATStep.H:22-307); settings:  \
     * chain_set(ACCEPTANCE_WINDOW | ACCEPTANCE_RIGIDITY), set_gaussian (sigma  \
     * itself, :121-133), set_uniform.  misc = trials, successes, last index,   \
     * indices left in the queue. */                                            \
    void* P##chain_create_vaat(int kind, int dim, uint64_t seed, uint32_t chain); \
    int P##chain_get_vaat(void* h, double* sigma, double* acceptance,         \
                          int32_t* acceptance_trials, int32_t* misc);         \
    void P##chain_destroy(void* h);                                           \
    int P##chain_set_fake(void* h, const orc_event* ev, long n,               \
                          const double* data150, double exposure);            \
    int P##chain_set_error_matrix(void* h, const double* e, int n);           \
    int P##chain_set(void* h, int field, double value);                       \
    int P##chain_set_gaussian(void* h, int d, double sigma);                  \
    int P##chain_set_uniform(void* h, int d, double lo, double hi);           \
    int P##chain_set_correlation(void* h, int d1, int d2, double c);          \
    int P##chain_start(void* h, const double* x0);                            \
    int P##chain_step(void* h, int nsteps, int metropolis, int32_t* accepted, \
                      double* llh_accepted, double* llh_proposed, double* x,  \
                      double* sigma);                                         \
    int P##chain_update_proposal(void* h);                                    \
    int P##chain_reset_proposal(void* h);                                     \
    int P##chain_get_state(void* h, double* scalars, double* accepted,        \
                           double* center, double* cov, double* decomp);      \
    double P##chain_llh(void* h, const double* x);                            \
    int P##chain_fake_hist(void* h, const double* x, double* out150);         \
    /* the debugging modes of TProposeAdaptiveStep::operator() (:671-704) and        \
     * SetEstimatedCenter (:733-739) */                                             \
    int P##chain_force_step(void* h, const double* x);                            \
    int P##chain_set_scan(void* h, int dim);                                      \
    int P##chain_set_center(void* h, const double* v);                            \
    const char* P##last_error(void);

ORC_DECLARE(ref_)
ORC_DECLARE(orc_)

/* sMCMC::TSimpleHMC<L, G> (TSimpleHMC.H:119-973), one chain per handle, same
 * injected stream: step s consumes slots 0..n-1 (momentum refresh), n (epsilon
 * jitter) and n+1 (accept) in the order the reference calls gRandom
 * (TSimpleHMC.H:568, :297, :347).  with_gradient selects
 * TSimpleHMC<L, L> (user gradient) or TSimpleHMC<L> (finite differences). */
enum {
    ORC_HMC_ALPHA = 0,          /* SetAlpha        :175 */
    ORC_HMC_MEAN_EPSILON = 1,   /* SetMeanEpsilon  :181 */
    ORC_HMC_LEAPFROG = 2        /* SetLeapFrog     :190 */
};
enum {
    ORC_HS_ACCEPTANCE = 0, ORC_HS_MEAN_EPSILON, ORC_HS_LEAPFROG, ORC_HS_REVERSAL_LEN,
    ORC_HS_ACCEPTED_POTENTIAL, ORC_HS_PROPOSED_POTENTIAL, ORC_HS_CENTRAL_POTENTIAL,
    ORC_HS_POTENTIAL_COUNT, ORC_HS_GRADIENT_COUNT, ORC_HS_STEP_COUNT, ORC_HS_COV_TRIALS,
    ORC_HS_AVERAGE_TRIALS, ORC_HS_EST_COV_TRACE, ORC_HS_CUR_COV_TRACE, ORC_HS_ORBIT_LENGTH,
    ORC_HS_STEPS_REMAINING, ORC_HS_STEPS_SINCE_UPDATE, ORC_HS_COUNT
};
/* Reference build only: Restore(tree, randomize = true) (TSimpleMCMC.H:309-316) from the
 * tree of `source`; the uniforms of the walk are draws k = 0, 1, ... of (seed, chain,
 * step 0xffffffff).  out[0] = TotalSteps of the entry that was adopted. */
int ref_chain_restore_random(void* h, void* source, int32_t* out);
#define ORC_DECLARE_HMC(P)                                                            \
    void* P##hmc_create(int kind, int dim, int with_gradient, uint64_t seed, uint32_t chain); \
    void P##hmc_destroy(void* h);                                                     \
    int P##hmc_set_error_matrix(void* h, const double* e, int n);                     \
    int P##hmc_set(void* h, int field, double value);                                 \
    int P##hmc_start(void* h, const double* x0);                                      \
    int P##hmc_step(void* h, int nsteps, int gradient_type, double* potential,        \
                    double* x, double* epsilon, int32_t* leapfrog);                   \
    int P##hmc_get_state(void* h, double* scalars, double* accepted, double* momentum, \
                         double* central, double* average, double* covariance, double* error);
ORC_DECLARE_HMC(ref_)
ORC_DECLARE_HMC(orc_)

#ifdef __cplusplus
}
#endif
#endif
