// vaat.cuh -- sMCMC::TProposeVAATStep (TProposeVAATStep.H:22-307), the adaptive
// variable-at-a-time proposal, for E chains: one THREAD per chain (a step touches
// one coordinate and three per-dimension scalars of the chain).
//
// Draws of a step in the order the reference calls gRandom: when the index
// queue is empty, n uniforms for the shuffle (:186-189, slots 0..n-1), then one
// draw for the proposed coordinate (:60-78); the accept uniform of
// TSimpleMCMC::Step follows in the next slot (VaatState::acceptSlot, read by
// kAccept).
//
// HBM layout (chain-major): sigma, acceptance : double[E][n];
// acceptanceTrials, queue : int[E][n]; VaatState[E].  The sampler-level scalars
// (accepted / proposed likelihood, step RMS, counters, fLastValue, fTrials,
// fSuccesses, fAcceptanceRigidity) live in ChainScalars as for the adaptive
// proposal; ChainScalars::sigma holds GetSigma() (:166-173), the mean step size.
#pragma once
#include "proposal.cuh"

namespace smcmc {

struct VaatState {
    int lastIndex;      // fLastIndex       :275
    int queueSize;      // fNextIndex.size() :272
    int acceptSlot;     // slot of the Metropolis uniform of the step in flight
    int pad_;
};

struct VaatArrays {
    double* sigma;          // fSigma            :299
    double* acceptance;     // fAcceptance       :290
    int* acceptanceTrials;  // fAcceptanceTrials :293
    int* queue;             // fNextIndex        :272
    VaatState* st;
    int window;             // fAcceptanceWindow (an int, :281)
    double target;          // fTargetAcceptance :296 (0.44)
};

// The tail of TSimpleMCMC::Start (:258-275) + InitializeState (:193-209).
__global__ void kVaatInit(ChainArrays a, VaatArrays v, int chains, int n, int32_t* ok) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= chains) return;
    ChainScalars s = a.sc[c];
    s.llhCalls += 1;                                                       // :539
    const bool good = devIsFinite(s.propLlh) && !(s.propLlh < -0.999999E+10);   // :265-268
    if (ok) ok[c] = good ? 1 : 0;
    s.started = good ? 1 : 0;
    if (good) {
        s.accLlh = s.propLlh;                                              // :270
        if (!s.initialized) s.lastValue = s.accLlh;                        // TProposeVAATStep.H:197-204 (first Start only)
        s.initialized = 1;
        double t = 0.0;
        for (int i = 0; i < n; ++i) t = __dadd_rn(t, v.sigma[(size_t)c * n + i]);
        s.sigma = __ddiv_rn(t, (double)n);
    }
    a.sc[c] = s;
}

// The head of TSimpleMCMC::Step (:376-406) with TProposeVAATStep::operator()
// (:40-80).  The functor starts from a copy of the accepted point (:52) and moves
// ONE coordinate.  The engine copies xAcc into xProp once (Start / Restore); from
// then on xProp differs from xAcc at most in the coordinate of the previous step
// (when that step was rejected), which this kernel puts back before it moves the
// next one -- 8 bytes per chain instead of a copy of the whole ensemble per step.
__global__ void __launch_bounds__(128)
kVaatPropose(ChainArrays a, VaatArrays v, PropSettings ps, int chains, uint64_t seed, uint32_t chainOffset,
             StepRef stepRef) {
    const uint32_t step = stepRef.get();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= chains) return;
    const int n = ps.n;
    ChainScalars s = a.sc[c];
    if (!s.started || s.status != 0) return;
    VaatState st = v.st[c];
    double* sigma = v.sigma + (size_t)c * n;
    double* acc = v.acceptance + (size_t)c * n;
    int* accTrials = v.acceptanceTrials + (size_t)c * n;
    int* queue = v.queue + (size_t)c * n;
    const uint32_t gchain = chainOffset + (uint32_t)c;
    uint32_t slot = 0;

    s.totalSteps += 1;                                                     // :376
    if (st.lastIndex >= 0)                                                 // :52, see above
        a.xProp[(size_t)c * n + st.lastIndex] = a.xAcc[(size_t)c * n + st.lastIndex];

    // ---- UpdateState, TProposeVAATStep.H:216-254 ---------------------------
    const double value = s.accLlh;
    s.trials += 1;
    const bool accepted = value != s.lastValue;                            // :223
    if (accepted) s.successes += 1;
    s.lastValue = value;
    if (st.lastIndex >= 0) {
        const int li = st.lastIndex;
        const int t = accTrials[li] + 1;                                   // :235
        accTrials[li] = t;
        const double m = __dmul_rn(1.0, (double)min(v.window, t));         // :236-240
        double av = __dmul_rn(acc[li], m);
        if (accepted) av = __dadd_rn(av, 1.0);
        av = __ddiv_rn(av, __dadd_rn(1.0, m));
        acc[li] = av;
        if ((double)t > __dmul_rn(0.1, (double)v.window) && s.rigidity > 0 && s.rigidity < 100.0) {   // :242-252
            const double ex = fmin(__ddiv_rn(1.0, 500.0), __ddiv_rn(1.0, __dmul_rn(s.rigidity, (double)v.window)));
            const double old = sigma[li];
            const double nv = fmax(__dmul_rn(old, pow(__ddiv_rn(av, v.target), ex)), 1.0E-4);
            sigma[li] = nv;
            // GetSigma(): the mean of fSigma, summed in index order (:166-173)
            double tot = 0.0;
            for (int i = 0; i < n; ++i) tot = __dadd_rn(tot, sigma[i]);
            s.sigma = __ddiv_rn(tot, (double)n);
        }
    }

    // ---- UpdateProposal, :176-190: refill and shuffle the index queue --------
    if (st.queueSize == 0) {
        st.lastIndex = -1;
        for (int i = 0; i < n; ++i) queue[i] = i;
        for (int i = 0; i < n; ++i) {
            const double u = __dmul_rn(1.0, smcmc_uniform(seed, gchain, step, slot++, SMCMC_STREAM_STEP));
            int sw = (int)(unsigned long long)__dmul_rn((double)n, u);
            if (sw >= n) sw = n - 1;              // u < 1, but n*u can round up to n
            const int tmp = queue[i];
            queue[i] = queue[sw];
            queue[sw] = tmp;
        }
        st.queueSize = n;
    }
    const int li = queue[st.queueSize - 1];                                // :58-59
    st.lastIndex = li;
    st.queueSize -= 1;

    // ---- the proposal for coordinate li, :60-78 ---------------------------------
    const double cur = a.xAcc[(size_t)c * n + li];
    double prop;
    if (ps.type[li] == 1) {
        const double u = smcmc_uniform(seed, gchain, step, slot++, SMCMC_STREAM_STEP);
        prop = __dadd_rn(ps.param1[li], __dmul_rn(__dsub_rn(ps.param2[li], ps.param1[li]), u));
    } else {
        double ev = 1.0;
        if (ps.param1[li] > 0) ev = ps.param1[li];
        const double g = smcmc_normal(seed, gchain, step, slot++, SMCMC_STREAM_STEP);
        const double r = __dadd_rn(0.0, __dmul_rn(ev, g));                 // TRandom::Gaus(0, ev)
        prop = __dadd_rn(cur, __dmul_rn(sigma[li], r));
    }
    a.xProp[(size_t)c * n + li] = prop;
    st.acceptSlot = (int)slot;
    v.st[c] = st;

    // ---- step RMS, TSimpleMCMC.H:391-406: one coordinate moved -------------------
    if (ps.stepRMSWindow > 0) {
        const double d = __dsub_rn(prop, cur);
        // the other n-1 terms of the sum are (x - x)^2 = +0
        const double sqr = __dadd_rn(0.0, __dmul_rn(d, d));
        double ms = __dmul_rn(s.stepRMS, s.stepRMS);
        ms = __dmul_rn(ms, (double)s.stepRMSTrials);
        ms = __dadd_rn(ms, sqr);
        ms = __ddiv_rn(ms, __dadd_rn((double)s.stepRMSTrials, 1.0));
        s.stepRMSTrials = min(ps.stepRMSWindow, s.stepRMSTrials + 1);
        s.stepRMS = __dsqrt_rn(ms);
    }
    a.sc[c] = s;
}

}  // namespace smcmc
