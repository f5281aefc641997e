// The reference's documentation example (reference TSimpleMCMC.H:122-156),
// written against include/TSimpleMCMC.H.  Prints per-step records that
// tests/test_gpu_cpp_facade.py compares with the oracle.
//   argv[1] = "unit" | "fake" | "fake2" ; argv[2] = chains ; argv[3] = steps
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

int main(int argc, char** argv) {
    const char* kind = argc > 1 ? argv[1] : "unit";
    int chains = argc > 2 ? std::atoi(argv[2]) : 1;
    int steps = argc > 3 ? std::atoi(argv[3]) : 100;
    if (!std::strcmp(kind, "fake")) return Run<FakeLikelihood>(chains, steps, false);
    if (!std::strcmp(kind, "fake2")) return Run<FakeLikelihood2>(chains, steps, false);
    if (!std::strcmp(kind, "vaat")) return RunVaat(chains, steps);
    return Run<TUnitGaussLogLikelihood>(chains, steps, true);
}
