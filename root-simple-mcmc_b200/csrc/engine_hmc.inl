// engine_hmc.inl -- host side of the TSimpleHMC entry points (included by
// engine.cu after the engine struct).  Queues the kernels of hmc.cuh.

namespace {

constexpr size_t kHmcFdWorkBytes = 512u << 20;     // finite-difference work points per pass

void hmcAllocate(smcmc_engine* e) {
    HmcHost& h = e->hmc;
    if (h.allocated) return;
    const size_t E = e->E(), n = e->n(), tri = e->tri();
    h.qAcc.reserve(E * n);
    h.pAcc.reserve(E * n);
    h.qProp.reserve(E * n);
    h.pProp.reserve(E * n);
    h.p0.reserve(E * n);
    h.grad.reserve(E * n);
    h.central.reserve(E * n);
    h.average.reserve(E * n);
    h.repairedDiag.reserve(E * n);
    // One covariance estimate per chain (the reference, TSimpleHMC.H:665-695) or one for the ensemble
    // (hmc.cuh, kHmcPooled*).  Automatic: pooled when the per-chain triangles would take 256 MB or more
    // (n = 500: from 269 chains on) and nothing needs a private estimate (the covariant gradient reads
    // the chain's own error matrix).  SMCMC_HMC_POOLED_COVARIANCE / SMCMC_HMC_POOLED=0|1 decide otherwise.
    h.pooled = h.pooledSetting < 0 ? (E * tri * sizeof(double) >= (256u << 20) && !h.keepError) : h.pooledSetting != 0;
    if (const char* k = std::getenv("SMCMC_HMC_POOLED")) h.pooled = atoi(k) != 0;
    if (h.pooled && h.keepError)
        throw Error(SMCMC_ERR_LOGIC, "the covariant gradient (SMCMC_HMC_KEEP_ERROR_MATRIX) needs the per-chain covariance");
    h.exxtT.reserve(E);
    CUDA_CHECK(cudaMemset(h.exxtT.get(), 0xFF, E * sizeof(double)));
    if (h.pooled) {
        h.poolMask.reserve(E);
        h.poolStats.reserve(1 + n + tri);
        h.poolExxt.reserve(tri);
        h.poolAverage.reserve(n);
        h.poolDiag.reserve(n);
        h.poolScratch.reserve(2 * n * n + 4 * n);
        h.poolLlh.reserve(1);
        h.pool.reserve(1);
        CUDA_CHECK(cudaMemset(h.poolMask.get(), 0, h.poolMask.bytes()));
        h.deferK = 0;
    } else {
        h.exxt.reserve(E * tri);
        // fEXXT is rewritten once per deferK steps instead of every step when the triangles of the
        // ensemble are larger than the L2 can hold (SMCMC_HMC_DEFER=k forces k, 0 = every step)
        h.deferK = (E * tri * sizeof(double) >= (256u << 20)) ? 16 : 0;
        if (const char* k = std::getenv("SMCMC_HMC_DEFER")) h.deferK = std::max(0, std::min(64, atoi(k)));
    }
    if (h.deferK > 0) {
        h.ring.reserve(E * h.deferK * n);
        h.ringT.reserve(E * h.deferK);
        h.pending.reserve(E);
        h.exxtDiag.reserve(E * n);
        CUDA_CHECK(cudaMemset(h.pending.get(), 0, h.pending.bytes()));
        CUDA_CHECK(cudaMemset(h.exxtDiag.get(), 0, h.exxtDiag.bytes()));
        const size_t flushSmem = exxtFlushSmem((int)n, h.deferK);
        if (flushSmem > 48 * 1024)
            CUDA_CHECK(cudaFuncSetAttribute(kHmcExxtFlush, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)flushSmem));
    }
    h.llh.reserve(E);
    h.sc.reserve(E);
    h.leapSteps.reserve(E);
    h.updateList.reserve(E);
    h.counters.reserve(4);
    CUDA_CHECK(cudaMallocHost((void**)&h.hostCounters, 4 * sizeof(int)));
    std::vector<HmcScalars> init(E);
    std::memset(init.data(), 0, sizeof(HmcScalars) * E);
    for (size_t c = 0; c < E; ++c) init[c].leapFrogSteps = 10;            // TSimpleHMC.H:133
    CUDA_CHECK(cudaMemcpy(h.sc.get(), init.data(), sizeof(HmcScalars) * E, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemset(h.pAcc.get(), 0, h.pAcc.bytes()));
    CUDA_CHECK(cudaMemset(h.pProp.get(), 0, h.pProp.bytes()));
    CUDA_CHECK(cudaMemset(h.grad.get(), 0, h.grad.bytes()));
    CUDA_CHECK(cudaMemset(h.leapSteps.get(), 0, h.leapSteps.bytes()));
    h.allocated = true;
}

HmcArrays hmcArrays(smcmc_engine* e) {
    HmcHost& h = e->hmc;
    HmcArrays a;
    a.qAcc = h.qAcc.get();
    a.pAcc = h.pAcc.get();
    a.qProp = h.qProp.get();
    a.pProp = h.pProp.get();
    a.p0 = h.p0.get();
    a.grad = h.grad.get();
    a.gradCur = nullptr;                 // set by hmcStepOnce for the steps that keep the gradients
    a.gradEnd = nullptr;
    a.central = h.central.get();
    a.average = h.average.get();
    a.exxt = h.exxt.get();
    a.exxtT = h.exxtT.get();
    a.ring = h.ring.get();
    a.ringT = h.ringT.get();
    a.pending = h.pending.get();
    a.exxtDiag = h.exxtDiag.get();
    a.deferK = h.deferK;
    a.pooled = h.pooled ? 1 : 0;
    a.poolMask = h.poolMask.get();
    a.poolStats = h.poolStats.get();
    a.poolExxt = h.poolExxt.get();
    a.poolAverage = h.poolAverage.get();
    a.poolDiag = h.poolDiag.get();
    a.pool = h.pool.get();
    a.estErr = h.keepError ? h.estErr.get() : nullptr;
    a.repairedDiag = h.repairedDiag.get();
    a.sc = h.sc.get();
    a.leapSteps = h.leapSteps.get();
    a.counters = h.counters.get();
    a.updateList = h.updateList.get();
    a.eigScratch = e->eigScratch.get();
    a.eigLocks = e->eigLocks.get();
    a.eigSlots = e->eigSlots;
    return a;
}

void requireHmcStarted(smcmc_engine* e) {
    if (!e->hmc.started) throw Error(SMCMC_ERR_INVALID_ARGUMENT, "Must initialize starting point");   // TSimpleHMC.H:280-284
}

// Which gradient PotentialGradient(type) ends up computing (:467-532).
enum HmcGradientMode { kGradUser, kGradFinite, kGradCovariant, kGradZero };

HmcGradientMode hmcResolveGradient(smcmc_engine* e, int type) {
    const bool haveUser = e->hmc.userGradient &&
                          (e->cfg.likelihood == SMCMC_LLH_DUMMY || e->cfg.likelihood == SMCMC_LLH_HARD ||
                           (e->cfg.likelihood == SMCMC_LLH_USER && e->userOps.gradient));
    switch (type) {
    case 2:
        if (!e->hmc.keepError)
            throw Error(SMCMC_ERR_LOGIC, "gradient type 2 needs SMCMC_HMC_KEEP_ERROR_MATRIX set before smcmc_hmc_start");
        return kGradCovariant;
    case 3: return kGradFinite;
    case 4:
        if (!haveUser) throw Error(SMCMC_ERR_LOGIC, "gradient type 4 needs a user gradient");        // :521
        return kGradUser;
    case 5: return kGradZero;
    default: return haveUser ? kGradUser : kGradFinite;                                           // :472-507
    }
}

void hmcGradient(smcmc_engine* e, HmcGradientMode mode, int k) {
    HmcHost& h = e->hmc;
    HmcArrays a = hmcArrays(e);
    const int E = e->E(), n = e->n();
    switch (mode) {
    case kGradUser: {
        if (e->cfg.likelihood == SMCMC_LLH_USER) {
            const int rc = e->userOps.gradient(e->userOps.ctx, h.qProp.get(), E, n, h.grad.get(), h.leapSteps.get(), k,
                                               (void*)e->stream);
            if (rc != 0) throw Error(SMCMC_ERR_CUDA, std::string("user gradient launch failed: ") +
                                                     cudaGetErrorString((cudaError_t)rc));
            e->launched();
            break;
        }
        if (e->cfg.likelihood == SMCMC_LLH_HARD) {
            kHardGradient<<<ceilDiv((long long)E * n, 256), 256, 0, e->stream>>>(h.qProp.get(), h.grad.get(),
                                                                                h.leapSteps.get(), k, E, n);
            e->launched();
            break;
        }
        if (e->errDim != n) throw Error(SMCMC_ERR_LOGIC, "error matrix not set (smcmc_dummy_set_error)");
        if (e->dummyMode == SMCMC_DUMMY_TENSOR) {
            launchDummyContractDmma(e->stream, h.qProp.get(), e->errMatrix.get(), h.grad.get(), h.leapSteps.get(), k, E, n, 0);
        } else {
            dim3 grid(ceilDiv(n, kGemmBN), ceilDiv(E, kGemmBM));
            kDummyGradient<<<grid, 256, 0, e->stream>>>(h.qProp.get(), e->errMatrix.get(), h.grad.get(),
                                                        h.leapSteps.get(), k, E, n);
        }
        e->launched();
        break;
    }
    case kGradFinite: {
        const size_t perChain = (size_t)2 * n * n * sizeof(double);
        int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)E, kHmcFdWorkBytes / perChain));
        h.fdWork.reserve((size_t)chunk * 2 * n * n);
        h.fdLlh.reserve((size_t)chunk * 2 * n);
        for (int base = 0; base < E; base += chunk) {
            const int cc = std::min(chunk, E - base);
            const size_t total = (size_t)cc * 2 * n * n;
            kHmcFdPoints<<<ceilDiv((long long)total, 256), 256, 0, e->stream>>>(h.qProp.get(), h.fdWork.get(), base, cc, n);
            e->launched();
            e->evaluate(h.fdWork.get(), cc * 2 * n, h.fdLlh.get(), nullptr);
            kHmcFdGradient<<<ceilDiv((long long)cc * n, 256), 256, 0, e->stream>>>(h.fdLlh.get(), h.grad.get(),
                                                                                  h.leapSteps.get(), k, base, cc, n);
            e->launched();
        }
        break;
    }
    case kGradCovariant:
        kHmcCovariantGradient<<<ceilDiv(E, kWarpsPerBlock), kWarpsPerBlock * 32,
                                (size_t)kWarpsPerBlock * n * sizeof(double), e->stream>>>(a, n, E, k);
        e->launched();
        break;
    case kGradZero:
        if (k == 0) {
            kHmcZeroGradient<<<ceilDiv((long long)E * n, 256), 256, 0, e->stream>>>(a, n, E);
            e->launched();
        }
        break;
    }
}

int hmcReadCounter(smcmc_engine* e, int which) {
    HmcHost& h = e->hmc;
    // ([0], [2] and [3] are written by the same kernel and read together)
    const int first = which == 1 ? 1 : 0, count = which == 1 ? 1 : 4;
    CUDA_CHECK(cudaMemcpyAsync(h.hostCounters + first, h.counters.get() + first, count * sizeof(int),
                               cudaMemcpyDeviceToHost, e->stream));
    CUDA_CHECK(cudaStreamSynchronize(e->stream));
    return h.hostCounters[which];
}

// apply the recorded UpdateCovariance calls to fEXXT: of every chain (list == nullptr) or of
// the `count` chains of `list` (hmc.cuh, kHmcExxtFlush)
void hmcFlushExxt(smcmc_engine* e, const int* list, int count) {
    HmcHost& h = e->hmc;
    if (h.deferK <= 0 || count <= 0) return;
    HmcArrays a = hmcArrays(e);
    const int n = e->n();
    const long long tri = (long long)n * (n + 1) / 2;
    const int perZ = std::min(count, 32768);
    dim3 grid(ceilDiv(tri, kExxtPerBlock), perZ, ceilDiv(count, perZ));
    kHmcExxtFlush<<<grid, kExxtThreads, exxtFlushSmem(n, h.deferK), e->stream>>>(a, n, count, list);
    e->launched();
    kHmcExxtFlushDone<<<ceilDiv(count, 256), 256, 0, e->stream>>>(a, count, list);
    e->launched();
    if (!list) h.sinceFlush = 0;
}

void hmcStepOnce(smcmc_engine* e, int type, const HmcTraceDev& tr, int traceStep) {
    HmcHost& h = e->hmc;
    const HmcGradientMode mode = hmcResolveGradient(e, type);
    HmcArrays a = hmcArrays(e);
    const int E = e->E(), n = e->n();
    const int blocks = ceilDiv(E, kWarpsPerBlock), threads = kWarpsPerBlock * 32;
    const size_t smem = (size_t)kWarpsPerBlock * n * sizeof(double);
    if (h.alpha < 0.0) h.alpha = 0.0;                                     // :565
    CUDA_CHECK(cudaMemsetAsync(h.counters.get(), 0, 4 * sizeof(int), e->stream));
    kHmcBegin<<<blocks, threads, smem, e->stream>>>(a, n, E, h.alpha, e->cfg.seed, e->cfg.chain_offset, e->stepIndex);
    e->launched();
    // dense Gaussian on the tensor cores: gradient, kick and drift of a leap-frog stage in ONE launch
    // (contraction.cuh, kHmcLeapDmma); SMCMC_HMC_NO_FUSE=1 keeps gradient kernel + kHmcKickDrift
    const bool fused = mode == kGradUser && e->cfg.likelihood == SMCMC_LLH_DUMMY && e->dummyMode == SMCMC_DUMMY_TENSOR &&
                       e->errDim == n && !std::getenv("SMCMC_HMC_NO_FUSE");
    const int maxSteps = hmcReadCounter(e, 0);
    // the chains' trajectory lengths are fetched only when they differ (shortest < longest)
    const int minSteps = h.hostCounters[3] > 0 ? (1 << 20) - h.hostCounters[3] : maxSteps;
    const bool wantOrder = fused && !std::getenv("SMCMC_HMC_NO_ORDER") &&
                           (minSteps < maxSteps || std::getenv("SMCMC_HMC_ORDER_ALWAYS"));
    if (wantOrder && maxSteps >= 1) {
        if (!h.hostSteps) CUDA_CHECK(cudaMallocHost((void**)&h.hostSteps, (size_t)2 * E * sizeof(int)));
        CUDA_CHECK(cudaMemcpyAsync(h.hostSteps, h.leapSteps.get(), (size_t)E * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
        CUDA_CHECK(cudaStreamSynchronize(e->stream));
    }
    const int countPotentials = (mode == kGradFinite) ? 2 * n : 0;
    // ... and when every running chain has a trajectory, the potential at its end comes out of the
    // chain's last gradient launch (LeapFused::endPartial) instead of a GEMM of its own
    bool potentialDone = false;
    bool keepGradients = false;
    if (maxSteps >= 1 && fused) {
        const int colBlocks = ceilDiv(n, kDmmaBN);
        h.qAlt.reserve((size_t)E * n);
        h.uturn.reserve((size_t)2 * E * colBlocks);
        const bool endPotential = h.hostCounters[2] == 0 && !std::getenv("SMCMC_HMC_SEPARATE_POTENTIAL");
        if (endPotential) e->dummyPartials.reserve((size_t)E * colBlocks);
        // The gradient at the starting point is known from the previous step (kHmcLeapCached): every
        // step keeps the gradients at its start and at its end point, kHmcAccept the one at the point
        // the chain goes on from.  Needs every running chain to have a trajectory, this step and the last.
        keepGradients = h.hostCounters[2] == 0 && !std::getenv("SMCMC_HMC_NO_GRADIENT_CACHE");
        const bool cached = keepGradients && h.gradCacheReady && h.gradCacheMode == 100;
        if (keepGradients) {
            h.gradCur.reserve((size_t)E * n);
            h.gradEnd.reserve((size_t)E * n);
        }
        LeapFused f;
        f.p = h.pProp.get();
        f.p0 = h.p0.get();
        f.epsilon = &h.sc.get()->epsilon;
        f.okLeap = &h.sc.get()->okLeap;
        f.scalarStride = (int)(sizeof(HmcScalars) / sizeof(double));
        f.leapSteps = h.leapSteps.get();
        f.uturn = h.uturn.get();
        f.blocks = colBlocks;
        f.endPartial = endPotential ? e->dummyPartials.get() : nullptr;
        f.gradStart = keepGradients && !cached ? h.gradCur.get() : nullptr;
        f.gradEnd = keepGradients ? h.gradEnd.get() : nullptr;
        // Ragged trajectory lengths (each chain tunes its own, :302-323): the launches take the chains in
        // order of length, so that the row tiles behind the chains that still run skip the GEMM
        // (kHmcLeapDmma).  The lengths are read back next to the counter above; ordering them is a
        // counting sort on the host.  Not worth it when (nearly) all tiles stay busy to the end.
        std::vector<int> gemmTiles;
        f.order = nullptr;
        f.gemmTiles = 0;
        if (wantOrder) {
            // (hmc_order.h: counting sort -- longest first, equal lengths in chain order, chains without a
            // trajectory last -- and the row tiles each launch needs; checked on the CPU by tests/test_rng.py)
            std::vector<int> scratch(maxSteps + 2);
            gemmTiles.resize(maxSteps + 1);
            const long long allTiles = (long long)(maxSteps + 1) * ceilDiv(E, kDmmaBM);
            const long long busyTiles = smcmc_hmc_order(h.hostSteps, E, maxSteps, kDmmaBM, nullptr, gemmTiles.data(), scratch.data());
            if (busyTiles * 100 <= allTiles * 97 || std::getenv("SMCMC_HMC_ORDER_ALWAYS")) {
                int* order = h.hostSteps + E;
                smcmc_hmc_order(h.hostSteps, E, maxSteps, kDmmaBM, order, gemmTiles.data(), scratch.data());
                h.order.reserve(E);
                CUDA_CHECK(cudaMemcpyAsync(h.order.get(), order, (size_t)E * sizeof(int), cudaMemcpyHostToDevice, e->stream));
                f.order = h.order.get();
            }
        }
        for (int k = 0; k <= maxSteps; ++k) {
            f.qIn = h.qProp.get();
            f.qOut = h.qAlt.get();
            if (f.order) f.gemmTiles = gemmTiles[k];
            if (k == 0 && cached) kHmcLeapCached<<<ceilDiv(E, 4), 128, 0, e->stream>>>(f, h.gradCur.get(), E, n);
            else launchHmcLeapDmma(e->stream, e->errMatrix.get(), f, k, E, n);
            e->launched();
            h.qProp.swap(h.qAlt);                                        // the proposed positions are in the buffer just written
        }
        a = hmcArrays(e);
        if (keepGradients) {
            a.gradCur = h.gradCur.get();
            a.gradEnd = h.gradEnd.get();
            h.gradCacheMode = 100;                                       // the fused tensor path
        }
        if (endPotential) {
            kDummyLlhFromPartials<<<ceilDiv(E, 128), 128, 0, e->stream>>>(e->dummyPartials.get(), colBlocks, E, h.llh.get());
            e->launched();
            potentialDone = true;
        }
    } else if (maxSteps >= 1) {
        // The same in the general path, for the gradients that are functions of the point alone -- the
        // built-in analytic ones and finite differences of a built-in likelihood (2n likelihoods per
        // gradient): the gradient kernel of k = 0 is skipped when the previous step left the gradient at
        // the chain's point.  (A user functor is called as often as the reference calls it.)
        const bool pure = e->cfg.likelihood != SMCMC_LLH_USER && (mode == kGradUser || mode == kGradFinite);
        keepGradients = pure && h.hostCounters[2] == 0 && !std::getenv("SMCMC_HMC_NO_GRADIENT_CACHE");
        const bool cached = keepGradients && h.gradCacheReady && h.gradCacheMode == (int)mode;
        if (keepGradients) {
            h.gradCur.reserve((size_t)E * n);
            h.gradEnd.reserve((size_t)E * n);
        }
        for (int k = 0; k <= maxSteps; ++k) {
            HmcArrays ak = a;
            if (k == 0 && cached) ak.grad = h.gradCur.get();
            else hmcGradient(e, mode, k);
            kHmcKickDrift<<<blocks, threads, smem, e->stream>>>(ak, n, E, k, countPotentials,
                                                                keepGradients && !cached ? h.gradCur.get() : nullptr,
                                                                keepGradients ? h.gradEnd.get() : nullptr);
            e->launched();
        }
        if (keepGradients) {
            a.gradCur = h.gradCur.get();
            a.gradEnd = h.gradEnd.get();
            h.gradCacheMode = (int)mode;
        }
    }
    if (!potentialDone) e->evaluate(h.qProp.get(), E, h.llh.get(), nullptr);   // :327
    kHmcPost<<<blocks, threads, smem, e->stream>>>(a, n, E, h.llh.get(), 1000000.0 /* fCovarianceWindow :134 */,
                                                   fused && maxSteps >= 1 ? 1 : 0);
    e->launched();
    if (h.pooled) {
        // UpdateCovariance on the pooled estimate: this step's sums over the marked chains, folded
        // into the running averages; then the trigger of UpdateErrorMatrix
        const long long tri = (long long)n * (n + 1) / 2;
        CUDA_CHECK(cudaMemsetAsync(h.poolStats.get(), 0, (size_t)(1 + n + tri) * sizeof(double), e->stream));
        e->poolAccumulateOn(h.qAcc.get(), nullptr, h.poolMask.get(), h.poolStats.get());
        kHmcPooledFold<<<ceilDiv(n + tri, 256), 256, 0, e->stream>>>(a, n);
        e->launched();
        kHmcPooledTrigger<<<1, 256, 0, e->stream>>>(a, n, 1000000.0 * (double)E);
        e->launched();
    } else if (h.deferK > 0) {
        if (++h.sinceFlush >= h.deferK) hmcFlushExxt(e, nullptr, E);     // :678-686, deferK steps at once
    } else {
        const long long tri = (long long)n * (n + 1) / 2;
        const int perZ = std::min(E, 32768);
        dim3 grid(ceilDiv(tri, kExxtPerBlock), perZ, ceilDiv(E, perZ));
        kHmcExxtUpdate<<<grid, kExxtThreads, (size_t)n * sizeof(double), e->stream>>>(a, n, E);   // :678-686
        e->launched();
    }
    const int updates = hmcReadCounter(e, 1);
    if (updates > 0 && h.pooled) {
        e->evaluate(h.poolAverage.get(), 1, h.poolLlh.get(), nullptr);    // :729, once for the ensemble
        kHmcPooledSpectrum<<<1, kHmcSpectrumThreads, 0, e->stream>>>(a, n, h.poolLlh.get(), h.poolScratch.get());
        e->launched();
        kHmcPooledApply<<<ceilDiv(E, 128), 128, 0, e->stream>>>(a, n, E);
        e->launched();
    } else if (updates > 0) {
        hmcFlushExxt(e, h.updateList.get(), updates);                     // UpdateErrorMatrix reads fEXXT
        h.avgPts.reserve((size_t)updates * n);
        h.avgLlh.reserve(updates);
        kHmcGatherAverage<<<ceilDiv((long long)updates * n, 256), 256, 0, e->stream>>>(a, n, updates, h.avgPts.get());
        e->launched();
        e->evaluate(h.avgPts.get(), updates, h.avgLlh.get(), nullptr);    // :729
        kHmcErrorMatrix<<<ceilDiv(updates, kWarpsPerBlock), threads, 0, e->stream>>>(a, n, updates, h.avgLlh.get());
        e->launched();
    }
    h.gradCacheReady = keepGradients;                                     // (kHmcAccept below completes it)
    const uint32_t acceptSlot = (h.alpha >= 1.0 ? 0u : (uint32_t)n) + 1u;
    kHmcAccept<<<blocks, threads, 0, e->stream>>>(a, n, E, e->cfg.seed, e->cfg.chain_offset, e->stepIndex,
                                                  acceptSlot, tr, traceStep);
    e->launched();
    ++e->stepIndex;
}

}  // namespace

extern "C" {

int smcmc_hmc_set(smcmc_engine* e, int setting, double v) {
    return guarded(e, [&]() {
        HmcHost& h = e->hmc;
        switch (setting) {
        case SMCMC_HMC_ALPHA: h.alpha = v; return;
        case SMCMC_HMC_USER_GRADIENT:
            if (v != 0.0 && e->cfg.likelihood != SMCMC_LLH_DUMMY && e->cfg.likelihood != SMCMC_LLH_HARD &&
                !(e->cfg.likelihood == SMCMC_LLH_USER && e->userOps.gradient))
                throw Error(SMCMC_ERR_INVALID_ARGUMENT, "only TDummyLogLikelihood, THardLogLikelihood and user functors with a "
                                                        "gradient entry provide a gradient functor");
            h.userGradient = (v != 0.0);
            return;
        case SMCMC_HMC_POOLED_COVARIANCE:
            if (h.allocated) throw Error(SMCMC_ERR_LOGIC, "set SMCMC_HMC_POOLED_COVARIANCE before the first HMC call that allocates state");
            h.pooledSetting = v < 0 ? -1 : (v != 0.0);
            return;
        case SMCMC_HMC_KEEP_ERROR_MATRIX:
            if (h.started) throw Error(SMCMC_ERR_LOGIC, "set SMCMC_HMC_KEEP_ERROR_MATRIX before smcmc_hmc_start");
            h.keepError = (v != 0.0);
            return;
        case SMCMC_HMC_MEAN_EPSILON:
        case SMCMC_HMC_LEAPFROG:
            hmcAllocate(e);
            kHmcSetScalar<<<ceilDiv(e->E(), 128), 128, 0, e->stream>>>(h.sc.get(), e->E(), setting, v);
            e->launched();
            return;
        default: throw Error(SMCMC_ERR_INVALID_ARGUMENT, "unknown HMC setting");
        }
    });
}

int smcmc_hmc_start(smcmc_engine* e, const double* x0) {
    return guarded(e, [&]() {
        if (!x0) throw Error(SMCMC_ERR_INVALID_ARGUMENT, "null starting points");
        HmcHost& h = e->hmc;
        hmcAllocate(e);
        const size_t E = e->E(), n = e->n();
        if (h.keepError) h.estErr.reserve(E * n * n);
        // scratch for UpdateErrorMatrix: eigenvalues and the inverse share the engine's slots
        CUDA_CHECK(cudaMemcpyAsync(h.qAcc.get(), x0, E * n * sizeof(double), cudaMemcpyHostToDevice, e->stream));
        h.gradCacheReady = false;                                          // new points: no gradient is known there
        e->evaluate(h.qAcc.get(), (int)E, h.llh.get(), nullptr);           // SetPosition, :221
        h.sinceFlush = 0;                                                  // kHmcStart empties the rings
        if (h.pooled) {
            HmcPooled init;
            std::memset(&init, 0, sizeof init);
            init.estCovTrace = (double)n;                                  // :263
            CUDA_CHECK(cudaMemcpyAsync(h.pool.get(), &init, sizeof init, cudaMemcpyHostToDevice, e->stream));
            CUDA_CHECK(cudaMemsetAsync(h.poolExxt.get(), 0, h.poolExxt.bytes(), e->stream));
            CUDA_CHECK(cudaMemsetAsync(h.poolAverage.get(), 0, h.poolAverage.bytes(), e->stream));
            CUDA_CHECK(cudaStreamSynchronize(e->stream));                  // `init` leaves scope
        }
        kHmcStart<<<ceilDiv((long long)E, kWarpsPerBlock), kWarpsPerBlock * 32, 0, e->stream>>>(
            hmcArrays(e), (int)n, (int)E, h.llh.get(), h.firstStart ? 1 : 0);
        e->launched();
        CUDA_CHECK(cudaStreamSynchronize(e->stream));
        h.firstStart = false;
        h.started = true;
        if ((size_t)kWarpsPerBlock * n * sizeof(double) > 48 * 1024) {
            const int bytes = (int)((size_t)kWarpsPerBlock * n * sizeof(double));
            CUDA_CHECK(cudaFuncSetAttribute(kHmcBegin, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
            CUDA_CHECK(cudaFuncSetAttribute(kHmcKickDrift, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
            CUDA_CHECK(cudaFuncSetAttribute(kHmcPost, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
            CUDA_CHECK(cudaFuncSetAttribute(kHmcCovariantGradient, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        }
    });
}

int smcmc_hmc_set_position(smcmc_engine* e, const double* x) {
    return guarded(e, [&]() {
        requireHmcStarted(e);
        if (!x) throw Error(SMCMC_ERR_INVALID_ARGUMENT, "null points");
        HmcHost& h = e->hmc;
        const size_t E = e->E(), n = e->n();
        CUDA_CHECK(cudaMemcpyAsync(h.qAcc.get(), x, E * n * sizeof(double), cudaMemcpyHostToDevice, e->stream));
        h.gradCacheReady = false;
        e->evaluate(h.qAcc.get(), (int)E, h.llh.get(), nullptr);
        kHmcSetPosition<<<ceilDiv((long long)E, 128), 128, 0, e->stream>>>(hmcArrays(e), (int)E, h.llh.get());
        e->launched();
        CUDA_CHECK(cudaStreamSynchronize(e->stream));
    });
}

int smcmc_hmc_step(smcmc_engine* e, int nsteps, int type) {
    return guarded(e, [&]() {
        requireHmcStarted(e);
        HmcTraceDev none = {nullptr, nullptr, nullptr, nullptr, nullptr};
        for (int s = 0; s < nsteps; ++s) hmcStepOnce(e, type, none, -1);
    });
}

int smcmc_hmc_step_trace(smcmc_engine* e, int nsteps, int type, const smcmc_hmc_trace* trace) {
    return guarded(e, [&]() {
        requireHmcStarted(e);
        if (nsteps < 1) return;
        const size_t rows = (size_t)nsteps * e->E();
        DeviceBuffer<double> pot, pts, eps;
        DeviceBuffer<int32_t> lf, acc;
        HmcTraceDev tr = {nullptr, nullptr, nullptr, nullptr, nullptr};
        if (trace) {
            if (trace->potential) { pot.reserve(rows); tr.potential = pot.get(); }
            if (trace->points) { pts.reserve(rows * e->n()); tr.points = pts.get(); }
            if (trace->mean_epsilon) { eps.reserve(rows); tr.meanEpsilon = eps.get(); }
            if (trace->leapfrog) { lf.reserve(rows); tr.leapfrog = lf.get(); }
            if (trace->accepted) { acc.reserve(rows); tr.accepted = acc.get(); }
        }
        for (int s = 0; s < nsteps; ++s) hmcStepOnce(e, type, tr, s);
        if (trace) {
            if (trace->potential) CUDA_CHECK(cudaMemcpyAsync(trace->potential, pot.get(), rows * 8, cudaMemcpyDeviceToHost, e->stream));
            if (trace->points) CUDA_CHECK(cudaMemcpyAsync(trace->points, pts.get(), rows * e->n() * 8, cudaMemcpyDeviceToHost, e->stream));
            if (trace->mean_epsilon) CUDA_CHECK(cudaMemcpyAsync(trace->mean_epsilon, eps.get(), rows * 8, cudaMemcpyDeviceToHost, e->stream));
            if (trace->leapfrog) CUDA_CHECK(cudaMemcpyAsync(trace->leapfrog, lf.get(), rows * 4, cudaMemcpyDeviceToHost, e->stream));
            if (trace->accepted) CUDA_CHECK(cudaMemcpyAsync(trace->accepted, acc.get(), rows * 4, cudaMemcpyDeviceToHost, e->stream));
        }
        CUDA_CHECK(cudaStreamSynchronize(e->stream));
    });
}

int smcmc_hmc_get(smcmc_engine* e, int field, void* dst, size_t bytes) {
    return guarded(e, [&]() {
        if (!dst) throw Error(SMCMC_ERR_INVALID_ARGUMENT, "null destination");
        HmcHost& h = e->hmc;
        if (!h.allocated) throw Error(SMCMC_ERR_LOGIC, "HMC state does not exist yet (smcmc_hmc_start)");
        const size_t E = e->E(), n = e->n(), tri = e->tri();
        auto need = [&](size_t want) {
            if (bytes < want) throw Error(SMCMC_ERR_INVALID_ARGUMENT, "destination too small");
        };
        auto copyArray = [&](const void* src, size_t want) {
            need(want);
            CUDA_CHECK(cudaMemcpyAsync(dst, src, want, cudaMemcpyDeviceToHost, e->stream));
            CUDA_CHECK(cudaStreamSynchronize(e->stream));
        };
        if (field == SMCMC_HMC_F_POOLED_COVARIANCE || field == SMCMC_HMC_F_POOLED_AVERAGE || field == SMCMC_HMC_F_POOLED_SCALARS) {
            if (!h.pooled) throw Error(SMCMC_ERR_LOGIC, "the covariance is kept per chain (SMCMC_HMC_POOLED_COVARIANCE)");
            std::vector<double> ex(tri), avg(n);
            HmcPooled hp;
            CUDA_CHECK(cudaStreamSynchronize(e->stream));
            CUDA_CHECK(cudaMemcpy(ex.data(), h.poolExxt.get(), tri * 8, cudaMemcpyDeviceToHost));
            CUDA_CHECK(cudaMemcpy(avg.data(), h.poolAverage.get(), n * 8, cudaMemcpyDeviceToHost));
            CUDA_CHECK(cudaMemcpy(&hp, h.pool.get(), sizeof hp, cudaMemcpyDeviceToHost));
            double* out = (double*)dst;
            if (field == SMCMC_HMC_F_POOLED_AVERAGE) {
                need(n * 8);
                std::copy(avg.begin(), avg.end(), out);
            } else if (field == SMCMC_HMC_F_POOLED_COVARIANCE) {
                need(n * n * 8);
                for (size_t i = 0; i < n; ++i)
                    for (size_t j = 0; j <= i; ++j)
                        out[i * n + j] = out[j * n + i] = ex[i * (i + 1) / 2 + j] - avg[i] * avg[j];      // :688-689
            } else {
                need(SMCMC_HMC_POOLED_SCALAR_COUNT * 8);
                out[SMCMC_HMC_PS_TRIALS] = hp.trials;
                out[SMCMC_HMC_PS_EST_COV_TRACE] = hp.estCovTrace;
                out[SMCMC_HMC_PS_CUR_COV_TRACE] = hp.curCovTrace;
                out[SMCMC_HMC_PS_ORBIT_LENGTH] = hp.orbitLength;
                out[SMCMC_HMC_PS_MAX_SCALE] = hp.maxScale;
                out[SMCMC_HMC_PS_MIN_SCALE] = hp.minScale;
                out[SMCMC_HMC_PS_STEP_COUNT] = hp.stepCount;
                out[SMCMC_HMC_PS_UPDATES] = hp.updates;
                out[SMCMC_HMC_PS_REPAIRED] = hp.repaired;
            }
            return;
        }
        switch (field) {
        case SMCMC_HMC_F_ACCEPTED: copyArray(h.qAcc.get(), E * n * 8); return;
        case SMCMC_HMC_F_MOMENTUM: copyArray(h.pAcc.get(), E * n * 8); return;
        case SMCMC_HMC_F_PROPOSED: copyArray(h.qProp.get(), E * n * 8); return;
        case SMCMC_HMC_F_CENTRAL: copyArray(h.central.get(), E * n * 8); return;
        case SMCMC_HMC_F_AVERAGE: copyArray(h.average.get(), E * n * 8); return;
        case SMCMC_HMC_F_ERROR_MATRIX:
            if (!h.keepError || !h.estErr.count()) throw Error(SMCMC_ERR_LOGIC, "the error matrix is not kept (SMCMC_HMC_KEEP_ERROR_MATRIX)");
            copyArray(h.estErr.get(), E * n * n * 8);
            return;
        default: break;
        }
        std::vector<HmcScalars> sc(E);
        CUDA_CHECK(cudaMemcpyAsync(sc.data(), h.sc.get(), sizeof(HmcScalars) * E, cudaMemcpyDeviceToHost, e->stream));
        CUDA_CHECK(cudaStreamSynchronize(e->stream));
        if (field == SMCMC_HMC_F_SCALARS) {
            need(E * SMCMC_HMC_SCALAR_COUNT * 8);
            double* out = (double*)dst;
            for (size_t c = 0; c < E; ++c) {
                double* o = out + c * SMCMC_HMC_SCALAR_COUNT;
                const HmcScalars& s = sc[c];
                o[SMCMC_HMC_S_ACCEPTANCE] = s.acceptance;
                o[SMCMC_HMC_S_MEAN_EPSILON] = s.meanEpsilon;
                o[SMCMC_HMC_S_LEAPFROG] = s.leapFrogSteps;
                o[SMCMC_HMC_S_REVERSAL_LEN] = s.reversalLen;
                o[SMCMC_HMC_S_ACCEPTED_POTENTIAL] = s.accPotential;
                o[SMCMC_HMC_S_PROPOSED_POTENTIAL] = s.propPotential;
                o[SMCMC_HMC_S_CENTRAL_POTENTIAL] = s.centralPotential;
                o[SMCMC_HMC_S_POTENTIAL_COUNT] = s.potentialCount;
                o[SMCMC_HMC_S_GRADIENT_COUNT] = s.gradientCount;
                o[SMCMC_HMC_S_STEP_COUNT] = s.stepCount;
                o[SMCMC_HMC_S_COV_TRIALS] = s.covTrials;
                o[SMCMC_HMC_S_AVERAGE_TRIALS] = s.averageTrials;
                o[SMCMC_HMC_S_EST_COV_TRACE] = s.estCovTrace;
                o[SMCMC_HMC_S_CUR_COV_TRACE] = s.curCovTrace;
                o[SMCMC_HMC_S_ORBIT_LENGTH] = s.orbitLength;
                o[SMCMC_HMC_S_STEPS_REMAINING] = s.stepsRemaining;
                o[SMCMC_HMC_S_STEPS_SINCE_UPDATE] = s.stepsSinceUpdate;
            }
            return;
        }
        if (field == SMCMC_HMC_F_COVARIANCE && h.pooled)
            throw Error(SMCMC_ERR_LOGIC, "the covariance is pooled over the ensemble: read SMCMC_HMC_F_POOLED_COVARIANCE");
        if (field == SMCMC_HMC_F_COVARIANCE) {
            // fEstimatedCovariance = fEXXT - mean mean^T (:688-689), or what a
            // positive-definiteness repair left (:792-806); the identity before
            // the first step (:254-260).
            need(E * n * n * 8);
            std::vector<double> ex(E * tri), avg(E * n), diag(E * n);
            hmcFlushExxt(e, nullptr, (int)E);
            CUDA_CHECK(cudaStreamSynchronize(e->stream));
            CUDA_CHECK(cudaMemcpy(ex.data(), h.exxt.get(), ex.size() * 8, cudaMemcpyDeviceToHost));
            CUDA_CHECK(cudaMemcpy(avg.data(), h.average.get(), avg.size() * 8, cudaMemcpyDeviceToHost));
            CUDA_CHECK(cudaMemcpy(diag.data(), h.repairedDiag.get(), diag.size() * 8, cudaMemcpyDeviceToHost));
            double* out = (double*)dst;
            for (size_t c = 0; c < E; ++c) {
                double* o = out + c * n * n;
                const bool fresh = sc[c].covTrials == 0.0 && sc[c].averageTrials == 0.0;
                for (size_t i = 0; i < n; ++i)
                    for (size_t j = 0; j <= i; ++j) {
                        double v;
                        if (fresh) v = (i == j) ? 1.0 : 0.0;
                        else if (sc[c].repaired) v = (i == j) ? diag[c * n + i] : 0.0;
                        else {
                            volatile double prod = avg[c * n + i] * avg[c * n + j];     // no contraction
                            v = ex[c * tri + i * (i + 1) / 2 + j] - prod;
                        }
                        o[i * n + j] = o[j * n + i] = v;
                    }
            }
            return;
        }
        throw Error(SMCMC_ERR_INVALID_ARGUMENT, "unknown HMC field");
    });
}

}  // extern "C"
