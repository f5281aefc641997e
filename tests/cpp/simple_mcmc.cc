// The reference's documentation example (reference TSimpleMCMC.H:122-156),
// written against include/TSimpleMCMC.H.  Prints per-step records that
// tests/test_gpu_cpp_facade.py compares with the oracle.
//   argv[1] = "unit" | "fake" | "fake2" | "vaat" | "debug" | "restore_random" | "tree" ; argv[2] = chains ; argv[3] = steps
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "TSimpleMCMC.H"
#include "TProposeVAATStep.H"
#include "smcmc_likelihoods.H"

// Starting point: the origin, or the true event counts for example2
// (example2/FakeMCMC.C starts the chain at MCTrueValues).
template <class L>
static void StartingPoint(const L&, sMCMC::Vector&) {}
static void StartingPoint(const FakeLikelihood2& like, sMCMC::Vector& p) { p = like.MCTrueValues; }

template <class L>
static int Run(int chains, int steps, bool hints) {
    TTree tree("SimpleMCMC", "Tree of accepted points");
    sMCMC::TSimpleMCMC<L> mcmc(&tree, true);
    mcmc.SetChains(chains);
    mcmc.SetSeed(1);
    L& like = mcmc.GetLogLikelihood();
    like.Init();
    mcmc.GetProposeStep().SetDim(like.GetDim());
    if (hints) {
        mcmc.GetProposeStep().SetGaussian(3, 2.0);
        mcmc.GetProposeStep().SetUniform(4, -5, 5);
        mcmc.GetProposeStep().SetCorrelation(3, 4, 0.3);
    }
    sMCMC::Vector point(like.GetDim());
    StartingPoint(like, point);
    if (!mcmc.Start(point, false)) { std::printf("start failed\n"); return 2; }
    std::printf("start llh %.17g direct %.17g\n", mcmc.GetAcceptedLogLikelihood(), like(point));
    int accepted = 0;
    for (int i = 0; i < steps; ++i) {
        bool ok = mcmc.Step();
        accepted += ok;
        std::printf("step %d acc %d llh %.17g x0 %.17g sigma %.17g\n", i, (int)ok, mcmc.GetAcceptedLogLikelihood(),
                    mcmc.GetAccepted()[0], mcmc.GetProposeStep().GetSigma());
    }
    mcmc.GetProposeStep().UpdateProposal();
    mcmc.SaveStep();
    std::printf("entries %ld expected %d accepted %d calls %d trace %.17g\n", tree.GetEntries(),
                chains * (steps + 1), accepted, mcmc.GetLogLikelihoodCount(), mcmc.GetProposeStep().GetCovarianceTrace());
#ifndef SMCMC_HAVE_ROOT_TTREE
    const std::vector<std::vector<double> >& cov = tree.VectorColumn("AdaptiveCovariance");
    std::printf("last covariance size %zu first covariance size %zu\n", cov.back().size(), cov.front().size());
#endif
    try {
        sMCMC::TSimpleMCMC<L> unstarted;
        unstarted.Step();
        return 3;
    } catch (std::invalid_argument& e) {
        std::printf("caught invalid_argument: %s\n", e.what());
    }
    return 0;
}

// SimpleVAAT.C:31-66 written against the mirror: the variable-at-a-time proposal.
// Chain `chain` of seed 7 is the golden chain "vaat_unit5" of the reference build.
static int RunVaat(int chains, int steps) {
    TTree tree("SimpleMCMC", "Tree of accepted points");
    sMCMC::TSimpleMCMC<TUnitGaussLogLikelihood, sMCMC::TProposeVAATStep> mcmc(&tree, false);
    mcmc.SetChains(chains);
    mcmc.SetSeed(7);
    mcmc.SetDevice(0, 3);
    TUnitGaussLogLikelihood& like = mcmc.GetLogLikelihood();
    mcmc.GetProposeStep().SetDim(like.GetDim());
    sMCMC::Vector point(like.GetDim());
    if (!mcmc.Start(point, false)) { std::printf("start failed\n"); return 2; }
    int accepted = 0;
    for (int i = 0; i < steps; ++i) {
        bool ok = mcmc.Step();
        accepted += ok;
        std::printf("step %d acc %d llh %.17g x0 %.17g sigma %.17g\n", i, (int)ok, mcmc.GetAcceptedLogLikelihood(),
                    mcmc.GetAccepted()[0], mcmc.GetProposeStep().GetSigma());
    }
    std::printf("entries %ld accepted %d trials %d successes %d window %g\n", tree.GetEntries(), accepted,
                mcmc.GetProposeStep().GetTrials(), mcmc.GetProposeStep().GetSuccesses(),
                mcmc.GetProposeStep().GetAcceptanceWindow());
    return 0;
}

// The debugging modes (reference TSimpleMCMC.H:671-704, :733-739, :811-830) through the
// mirror: the call sequence of tests/helpers.py::run_debug_modes; chain 0 of the
// program is chain 5 of seed 61, the golden run "debug9" of the reference build.
static int RunDebug() {
    sMCMC::TSimpleMCMC<TUnitGaussLogLikelihood> mcmc;
    mcmc.SetSeed(61);
    mcmc.SetDevice(0, 5);
    TUnitGaussLogLikelihood& like = mcmc.GetLogLikelihood();
    like.SetDim(9);
    sMCMC::TProposeAdaptiveStep& prop = mcmc.GetProposeStep();
    prop.SetDim(9);
    prop.SetGaussian(3, 0.7);
    prop.SetUniform(6, -1.5, 2.0);
    if (!mcmc.Start(sMCMC::Vector(9, 0.2), false)) return 2;
    int index = 0;
    auto run = [&](int n, int metropolis) {
        for (int i = 0; i < n; ++i) {
            const bool ok = mcmc.Step(false, metropolis);
            std::printf("step %d acc %d llh %.17g x0 %.17g x3 %.17g x6 %.17g\n", index++, (int)ok,
                        mcmc.GetAcceptedLogLikelihood(), mcmc.GetAccepted()[0], mcmc.GetAccepted()[3], mcmc.GetAccepted()[6]);
        }
    };
    sMCMC::Vector forced(9), center(9);
    for (int i = 0; i < 9; ++i) {
        forced[i] = -0.4 + 0.1 * i;
        center[i] = 0.3 - 0.075 * i;
    }
    forced[4] = 0.0;                            // numpy.linspace(-0.4, 0.4, 9)[4] is exactly 0
    run(60, 0);
    prop.ForceStep(forced);
    run(1, 2);
    run(5, 0);
    if (!prop.SetEstimatedCenter(center)) return 3;
    prop.SetScanDimension(3);
    run(25, 0);
    prop.SetScanDimension(6);
    run(15, 0);
    prop.SetScanDimension(-1);
    prop.ForceStep(sMCMC::Vector(9, 0.05));
    run(1, 0);
    run(80, 0);
    std::printf("frozen %d calls %d\n", (int)prop.GetCovarianceFrozen(), mcmc.GetLogLikelihoodCount());
    try {
        prop.ForceStep(sMCMC::Vector(3, 0.0));
        return 4;
    } catch (std::invalid_argument& e) {
        std::printf("caught invalid_argument: %s\n", e.what());
    }
    return 0;
}

// Restore(tree, randomize = true) (reference :309-316): chains 0..5 of seed 71 each write
// 400 saved steps and the full state, then a new sampler restores a random entry.
static int RunRestoreRandom() {
    for (int chain = 0; chain < 6; ++chain) {
        TTree tree("SimpleMCMC", "Tree of accepted points");
        {
            sMCMC::TSimpleMCMC<TUnitGaussLogLikelihood> a(&tree);
            a.SetSeed(71);
            a.SetDevice(0, chain);
            a.GetLogLikelihood().SetDim(4);
            a.GetProposeStep().SetDim(4);
            if (!a.Start(sMCMC::Vector(4, 0.1), false)) return 2;
            for (int i = 0; i < 400; ++i) a.Step(true);
            a.SaveStep();
        }
        sMCMC::TSimpleMCMC<TUnitGaussLogLikelihood> b;
        b.SetSeed(71);
        b.SetDevice(0, chain);
        b.GetLogLikelihood().SetDim(4);
        b.GetProposeStep().SetDim(4);
        if (!b.Start(sMCMC::Vector(4, 0.0), false)) return 2;
        b.Restore(&tree, true);
        std::printf("pick %d x %.17g %.17g %.17g %.17g llh %.17g\n", chain, b.GetAccepted()[0], b.GetAccepted()[1],
                    b.GetAccepted()[2], b.GetAccepted()[3], b.GetAcceptedLogLikelihood());
    }
    return 0;
}

// "Accepted points written to the user's TTree" at ensemble size: E chains of the event
// likelihood, `steps` steps without the tree (StepMany) and `steps` steps through
// Step(true) -> SaveStep -> TTree::Fill per chain; prints the time per step of each.
#include <chrono>
static int RunTreeTiming(int chains, int steps, int dataEvents) {
    TTree tree("SimpleMCMC", "Tree of accepted points");
    sMCMC::TSimpleMCMC<FakeLikelihood> mcmc(&tree);
    mcmc.SetChains(chains);
    mcmc.SetSeed(3);
    FakeLikelihood& like = mcmc.GetLogLikelihood();
    like.Init(dataEvents, dataEvents, 10.0);      // 30 x dataEvents simulated events (FakeLikelihood.H:86-100)
    mcmc.GetProposeStep().SetDim(like.GetDim());
    if (!mcmc.Start(sMCMC::Vector(like.GetDim(), 0.0), false)) return 2;
    mcmc.StepMany(3);
    auto t0 = std::chrono::steady_clock::now();
    mcmc.StepMany(steps);
    auto t1 = std::chrono::steady_clock::now();
    for (int i = 0; i < steps; ++i) mcmc.Step(true);
    auto t2 = std::chrono::steady_clock::now();
    mcmc.SaveStep();
    auto t3 = std::chrono::steady_clock::now();
    const double plain = std::chrono::duration<double, std::milli>(t1 - t0).count() / steps;
    const double saved = std::chrono::duration<double, std::milli>(t2 - t1).count() / steps;
    std::printf("tree chains %d steps %d events %zu ms_per_step_plain %.6f ms_per_step_tree %.6f full_save_ms %.6f entries %ld\n",
                chains, steps, like.SimulatedSample.size(), plain, saved, std::chrono::duration<double, std::milli>(t3 - t2).count(),
                tree.GetEntries());
    return 0;
}

int main(int argc, char** argv) {
    const char* kind = argc > 1 ? argv[1] : "unit";
    int chains = argc > 2 ? std::atoi(argv[2]) : 1;
    int steps = argc > 3 ? std::atoi(argv[3]) : 100;
    if (!std::strcmp(kind, "fake")) return Run<FakeLikelihood>(chains, steps, false);
    if (!std::strcmp(kind, "fake2")) return Run<FakeLikelihood2>(chains, steps, false);
    if (!std::strcmp(kind, "vaat")) return RunVaat(chains, steps);
    if (!std::strcmp(kind, "debug")) return RunDebug();
    if (!std::strcmp(kind, "restore_random")) return RunRestoreRandom();
    if (!std::strcmp(kind, "tree")) return RunTreeTiming(chains, steps, argc > 4 ? std::atoi(argv[4]) : 1000);
    return Run<TUnitGaussLogLikelihood>(chains, steps, true);
}
