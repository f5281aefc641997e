// proposal_resident.cuh -- kStepsResident: `nsteps` WHOLE Metropolis steps
// (TSimpleMCMC::Step, TSimpleMCMC.H:370-496) of a chain in ONE launch, for the
// likelihoods that need nothing but the chain's own point (the analytic test
// likelihoods of simple_likelihoods.cuh).
//
// kProposeStaged (proposal_staged.cuh) is HBM-bound: per chain and step it reads
// and rewrites the packed covariance and reads the Cholesky factor, 31 KB at
// n = 50, and the likelihood and the accept rule are two more launches.  None of
// that traffic is needed when the likelihood is local to the chain: the chain's
// whole adaptive state (covariance, factor, current / central / last point)
// fits in the shared memory of one CTA, so the CTA loads it once (the same two
// bulk copies as kProposeStaged), runs the steps back to back out of shared
// memory -- proposal, likelihood, accept, UpdateState -- and writes the state
// back once.  HBM traffic per chain-step falls from 3 n(n+1)/2 doubles to
// (3 n(n+1)/2 + 5 n) / nsteps doubles; the step becomes bound by instruction issue
// (about 4.5 k warp-instructions per chain-step at n = 50, a third of them the
// 51 Philox / Box-Muller draws) and by the dependent chain of the scalars.
//
// Measured on B200 (profiles/r01_configs.jsonl): one chain, n = 5: 14.3 -> 2.9 us
// per step.  For an ensemble that fills the GPU several times over the gain is
// gone -- 65536 chains x 50 dims: 0.54 ms per step here, 0.50 ms for the
// three-launch step, both at about half of the issue slots -- so the engine uses
// this kernel for ensembles of up to one wave of CTAs (6 per SM) and the
// three-launch step beyond.  A one-warp-per-chain variant (no CTA barriers, 9
// chains per SM) was measured too: 0.57 ms, dropped.
//
// The arithmetic is kProposeStaged's, kSimpleLikelihood's and kAccept's,
// operation for operation and draw for draw (same Philox slots), so a chain
// advanced here is bit-identical to one advanced by the three-launch step
// (tests/test_gpu_resident.py).
//
// Roles: warp 0 carries the scalars (UpdateState :1721-1776), then -- after the
// proposal -- the step RMS (:391-406), the likelihood of the proposed point and
// the Metropolis rule (:410-495).  Warps 1-3 update the central point and the
// covariance (:1780-1820) and, while warp 0 evaluates the likelihood, draw the
// normals of the NEXT step (they depend on the step index only).
#pragma once
#include "proposal_staged.cuh"
#include "simple_likelihoods.cuh"

namespace smcmc {

struct ResidentShared {
    StagedShared st;
    double centerT;     // fCentralPointTrials / fCovarianceTrials seen by the workers
    double covT;
    int stop;           // the chain left the active state (status != 0)
    int pad_;
};

__host__ __device__ inline int residentChainBytes(int n, int covStride, int upkStride) {
    const int bytes = (covStride + upkStride + 6 * ((n + 1) & ~1)) * 8 + 16 + (int)sizeof(ResidentShared);
    return (bytes + 127) & ~127;
}

__device__ __forceinline__ double residentLikelihood(int kind, const double* x, int n, const double* __restrict__ errT) {
    switch (kind) {
    case SMCMC_LLH_UNIT_GAUSS: return llhUnitGauss(x, n);
    case SMCMC_LLH_DUMMY: {
        // llhDummy through the transposed matrix: errT[i][j] = Error(j, i)
        double s = 0.0;
        for (int i = 0; i < n; ++i) {
            const double hx = __dmul_rn(0.5, x[i]);
            const double* row = errT + (size_t)i * n;
#pragma unroll 4
            for (int j = 0; j < n; ++j) s = __dsub_rn(s, __dmul_rn(__dmul_rn(hx, __ldg(row + j)), x[j]));
        }
        return s;
    }
    case SMCMC_LLH_HORRIFIC: return llhHorrific(x, n);
    case SMCMC_LLH_HARD: return llhHard(x, n);
    default: return llhAsym(x, n);
    }
}

// the draws of one step (:709-724): r_i for the Gaussian dimensions, the uniform
// point for the others
__device__ __forceinline__ void residentDraw(const PropSettings& ps, double* zr, int n, uint64_t seed, uint32_t gchain,
                                             uint32_t step, int t, int workers) {
    // one draw per thread: the workers' draws of the next step run beside warp 0's likelihood, and for
    // a short chain-local likelihood their latency is the critical path (a Box-Muller pair per thread is
    // 1.5 x the latency of one branch)
    for (int i = t; i < n; i += workers) {
        if (ps.anyUniform && ps.type[i] == 1) {
            const double uu = smcmc_uniform(seed, gchain, step, (uint32_t)i, SMCMC_STREAM_STEP);
            zr[i] = __dadd_rn(ps.param1[i], __dmul_rn(__dsub_rn(ps.param2[i], ps.param1[i]), uu));
        } else {
            const double g = smcmc_normal(seed, gchain, step, (uint32_t)i, SMCMC_STREAM_STEP);
            zr[i] = __dadd_rn(0.0, __dmul_rn(1.0, g));              // TRandom::Gaus(0,1)
        }
    }
}

__global__ void __launch_bounds__(kStagedThreads, 6)
kStepsResident(ChainArrays a, PropSettings ps, int chains, uint64_t seed, uint32_t chainOffset, uint32_t step0,
               int nsteps, int metropolis, int llhKind, const double* __restrict__ errT) {
    extern __shared__ __align__(128) unsigned char stagedSmem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int c = blockIdx.x;
    const int n = ps.n;
    double* covS = reinterpret_cast<double*>(stagedSmem);
    double* uS = covS + ps.covStride;
    const int nE = (n + 1) & ~1;
    double* cur = uS + ps.upkStride;     // current (= accepted) point
    double* cen = cur + nE;              // central point
    double* dif = cen + nE;              // cur - cen, then the squared step per dimension
    double* zr = dif + nE;               // draws of the step
    double* lastS = zr + nE;             // fLastPoint
    double* prop = lastS + nE;           // proposed point
    uint64_t* bar = reinterpret_cast<uint64_t*>(prop + nE);
    ResidentShared* sh = reinterpret_cast<ResidentShared*>(bar + 2);

    ChainScalars* scp = a.sc + c;
    double* cov = a.cov + (size_t)c * ps.covStride;
    double* upk = a.upk + (size_t)c * ps.upkStride;
    const bool stageCov = !ps.covFrozen;
    const uint32_t covBytes = (uint32_t)ps.covStride * 8u, upkBytes = (uint32_t)ps.upkStride * 8u;
    if (tid == 0) {
        mbarInit(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbarExpectTx(bar, (stageCov ? covBytes : 0u) + upkBytes);
        if (stageCov) tmaLoad1D(covS, cov, covBytes, bar);
        tmaLoad1D(uS, upk, upkBytes, bar);
    }
    const bool active = scp->started && scp->status == 0;
    ChainScalars s = *scp;
    const uint32_t gchain = chainOffset + (uint32_t)c;
    constexpr int kWorkers = kStagedThreads - 32;
    for (int i = tid; i < n; i += kStagedThreads) {
        cur[i] = a.xAcc[(size_t)c * n + i];
        cen[i] = a.center[(size_t)c * n + i];
        lastS[i] = a.lastPoint[(size_t)c * n + i];
        prop[i] = a.xProp[(size_t)c * n + i];
    }
    if (tid == 0) {
        sh->centerT = s.centerTrials;
        sh->covT = s.covTrials;
        sh->stop = 0;
        sh->st.fromGlobal = !s.upperTri;
        sh->st.upperTri = s.upperTri;
    }
    if (warp > 0 && active) residentDraw(ps, zr, n, seed, gchain, step0, tid - 32, kWorkers);
    mbarWait(bar, 0);                   // every thread: the rows are visible to all of them
    __syncthreads();
    if (!active) return;

    int done = 0;
    for (; done < nsteps; ++done) {
        const uint32_t step = step0 + (uint32_t)done;
        double value = 0.0;
        if (warp == 0) {
            // ---- the scalars of UpdateState, :1721-1776, :1824-1826 ----------------
            value = s.accLlh;
            s.totalSteps += 1;                                          // :376
            const bool accepted = updateStateScalars(s, ps, value, cur[0], lastS[0]);
            s.centerTrials = fmin(ps.covWindow, __dadd_rn(s.centerTrials, 1.0));
            if (stageCov) s.covTrials = fmin(ps.covWindow, __dadd_rn(s.covTrials, 1.0));
            bool update = false;
            if (accepted) {
                s.nextUpdate -= 1;
                update = s.nextUpdate < 1;
            }
            if (lane == 0) {
                sh->st.sigma = s.sigma;
                sh->st.accepted = accepted;
                sh->st.update = update;
                sh->st.status = 0;
            }
        } else {
            // ---- central point (:1780-1788), covariance (:1795-1820) ---------------
            const int t = tid - 32;
            const double centerT = sh->centerT, centerT1 = __dadd_rn(centerT, 1.0);
            const double covT = sh->covT, covT1 = __dadd_rn(covT, 1.0);
            for (int i = t; i < n; i += kWorkers) {
                const double x = cur[i];
                double v = __dmul_rn(cen[i], centerT);
                v = __dadd_rn(v, x);
                v = __ddiv_rn(v, centerT1);
                cen[i] = v;
                dif[i] = __dsub_rn(x, v);
            }
            namedBarrier(1, kWorkers);               // dif[] complete
            if (stageCov) {
                const bool fast = covT1 >= 1.0 && covT1 <= 1152921504606846976.0;
                const double y = __ddiv_rn(1.0, covT1);
                const char* difB = reinterpret_cast<const char*>(dif);
                if (fast) {
#pragma unroll 2
                    for (int k = t; k < ps.tri; k += kWorkers) {
                        const uint32_t ij = __ldg(ps.ijTab + k);       // byte offsets of dif[i], dif[j]
                        const double r = __dmul_rn(*reinterpret_cast<const double*>(difB + (ij & 0xffffu)),
                                                   *reinterpret_cast<const double*>(difB + (ij >> 16)));
                        const double w = __dadd_rn(__dmul_rn(covS[k], covT), r);
                        covS[k] = divideByShared(w, covT1, y);
                    }
                } else {
                    for (int k = t; k < ps.tri; k += kWorkers) {
                        const uint32_t ij = __ldg(ps.ijTab + k);
                        const double r = __dmul_rn(*reinterpret_cast<const double*>(difB + (ij & 0xffffu)),
                                                   *reinterpret_cast<const double*>(difB + (ij >> 16)));
                        const double w = __dadd_rn(__dmul_rn(covS[k], covT), r);
                        covS[k] = __ddiv_rn(w, covT1);
                    }
                }
            }
        }
        __syncthreads();                                                // B
        if (sh->st.update) {                                            // rare: UpdateProposal
            // the global-memory code of proposal.cuh on warp 0: the covariance goes out,
            // the (possibly conditioned) covariance and the new factor come back
            if (stageCov)
                for (int k = tid; k < ps.covStride; k += kStagedThreads) cov[k] = covS[k];
            __threadfence();
            __syncthreads();
            if (warp == 0) {
                {
                    ChainScalars sl = s;
                    PropSettings psl = ps;
                    ChainArrays al = a;
                    warpUpdateProposal(sl, psl, al, cov, al.decomp + (size_t)c * n * n, cen, lastS, false, lane);
                    s = sl;
                }
                if (lane == 0) {
                    sh->st.sigma = s.sigma;
                    sh->st.status = s.status;
                    sh->st.fromGlobal = !s.upperTri;
                    sh->st.upperTri = s.upperTri;
                }
                __threadfence();
            }
            __syncthreads();
            if (stageCov)
                for (int k = tid; k < ps.covStride; k += kStagedThreads) covS[k] = __ldcg(cov + k);
            for (int k = tid; k < ps.upkStride; k += kStagedThreads) uS[k] = __ldcg(upk + k);
            __syncthreads();                                            // C
        }
        const int status = sh->st.status;
        if (status == 0) {
            // ---- the proposal, :709-724: x'_j = x_j + sum_i (fSigma r_i) U(i,j), i ascending
            const double sigma = sh->st.sigma;
            const bool upper = sh->st.upperTri != 0;
            const bool fromGlobal = sh->st.fromGlobal != 0;
            const double* u = a.decomp + (size_t)c * n * n;
            for (int j = tid; j < n; j += kStagedThreads) {
                double p = cur[j];
                if (!fromGlobal && !ps.anyUniform) {
                    const double* row = uS + j;
                    int len = nE;
                    int i = 0;
                    for (; i + 1 <= j; i += 2) {
                        const double2 z = *reinterpret_cast<const double2*>(zr + i);
                        p = __dadd_rn(p, __dmul_rn(__dmul_rn(sigma, z.x), row[0]));
                        p = __dadd_rn(p, __dmul_rn(__dmul_rn(sigma, z.y), row[len]));
                        row += 2 * len - 2;
                        len -= 2;
                    }
                    if (i == j) p = __dadd_rn(p, __dmul_rn(__dmul_rn(sigma, zr[i]), row[0]));
                } else if (ps.type[j] == 1) {
                    p = zr[j];
                } else {
                    const int iEnd = upper ? j + 1 : n;
                    for (int i = 0; i < iEnd; ++i) {
                        if (ps.type[i] == 1) continue;
                        const double uij = fromGlobal ? __ldcg(u + (size_t)i * n + j) : uS[upkOffset(i, nE) + j - (i & ~1)];
                        p = __dadd_rn(p, __dmul_rn(__dmul_rn(sigma, zr[i]), uij));
                    }
                }
                prop[j] = p;
                const double d = __dsub_rn(p, cur[j]);
                dif[j] = __dmul_rn(d, d);
            }
        }
        for (int i = tid; i < n; i += kStagedThreads) lastS[i] = cur[i];    // :1829-1830
        __syncthreads();                                                // D
        if (warp == 0) {
            s.lastValue = value;
            if (status == 0 && ps.stepRMSWindow > 0) {                      // :391-406
                double sqr = 0.0;
#pragma unroll 4
                for (int i = 0; i < n; ++i) sqr = __dadd_rn(sqr, dif[i]);
                double ms = __dmul_rn(s.stepRMS, s.stepRMS);
                ms = __dmul_rn(ms, (double)s.stepRMSTrials);
                ms = __dadd_rn(ms, sqr);
                ms = __ddiv_rn(ms, __dadd_rn((double)s.stepRMSTrials, 1.0));
                s.stepRMSTrials = min(ps.stepRMSWindow, s.stepRMSTrials + 1);
                s.stepRMS = __dsqrt_rn(ms);
            }
            if (status == 0) {
                // ---- likelihood of the proposed point and the Metropolis rule, :410-495
                s.llhCalls += 1;                                            // :539
                const double propLlh = residentLikelihood(llhKind, prop, n, errT);
                const double accLlh = s.accLlh;
                s.propLlh = propLlh;
                bool take;
                if (metropolis == 2) {                                      // :414-426
                    take = true;
                } else if (!devIsFinite(propLlh) || propLlh < -0.999999E+30) {  // :432-436
                    take = false;
                } else {
                    take = true;
                    const double delta = __dsub_rn(propLlh, accLlh);        // :441
                    if (delta < 0.0) {
                        if (metropolis == 1) take = false;                  // :448
                        else {
                            const double uu = __dmul_rn(1.0, smcmc_uniform(seed, gchain, step, (uint32_t)n, SMCMC_STREAM_STEP));
                            const double trial = log(uu);                   // :455
                            if (delta < trial) take = false;
                        }
                    }
                }
                if (take) {
                    s.accLlh = propLlh;                                     // :484
                    for (int i = lane; i < n; i += 32) cur[i] = prop[i];    // :485-491
                }
            }
            if (lane == 0) {
                sh->centerT = s.centerTrials;
                sh->covT = s.covTrials;
                sh->stop = status != 0;
            }
        } else if (status == 0 && done + 1 < nsteps) {
            // the draws of the next step do not depend on this step's outcome
            residentDraw(ps, zr, n, seed, gchain, step + 1u, tid - 32, kWorkers);
        }
        __syncthreads();                                                // E
        if (sh->stop) break;
    }

    // ---- the state goes back ---------------------------------------------------
    if (stageCov)
        for (int k = tid; k < ps.covStride; k += kStagedThreads) cov[k] = covS[k];
    for (int i = tid; i < n; i += kStagedThreads) {
        a.xAcc[(size_t)c * n + i] = cur[i];
        a.center[(size_t)c * n + i] = cen[i];
        a.lastPoint[(size_t)c * n + i] = lastS[i];
        a.xProp[(size_t)c * n + i] = prop[i];
    }
    if (tid == 0) *scp = s;
}

}  // namespace smcmc
