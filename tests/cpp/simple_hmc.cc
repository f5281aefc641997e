// The reference's SimpleHMC.C (reference SimpleHMC.C:15-84) written against
// include/TSimpleHMC.H, followed by a save / Restore round trip of the
// Metropolis sampler (reference SimpleMCMC.C:45-60, :151-157).
//   argv[1] = "hmc" | "restore" ; argv[2] = chains ; argv[3] = steps
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "TSimpleHMC.H"
#include "smcmc_likelihoods.H"

static int RunHmc(int chains, int steps) {
    TTree tree("SimpleHMC", "Tree of accepted points");
    sMCMC::TSimpleHMC<TDummyLogLikelihood, TDummyLogLikelihood> hmc(&tree);
    hmc.SetChains(chains);
    hmc.SetSeed(28);
    TDummyLogLikelihood& like = hmc.GetLogLikelihood();
    like.Init();
    sMCMC::Vector p(like.GetDim());
    for (std::size_t i = 0; i < p.size(); ++i) p[i] = 1.0;
    hmc.Start(p, true);
    for (int i = 0; i < steps; ++i) {
        hmc.Step(true);
        std::printf("step %d potential %.17g x0 %.17g eps %.17g leap %d\n", i, hmc.GetAcceptedPotential(),
                    hmc.GetAccepted()[0], hmc.GetMeanEpsilon(), (int)tree.IntColumn("Leapfrog").back());
    }
    std::printf("entries %ld expected %d calls %d gradients %d acceptance %.17g\n", tree.GetEntries(),
                chains * (steps + 1), hmc.GetPotentialCount(), hmc.GetGradientCount(), hmc.GetAcceptanceRate());
    try {
        sMCMC::TSimpleHMC<TUnitGaussLogLikelihood> unstarted;
        unstarted.Step();
        return 3;
    } catch (std::invalid_argument& e) {
        std::printf("caught invalid_argument: %s\n", e.what());
    }
    return 0;
}

static int RunRestore(int chains, int steps) {
    // first run: writes the chain and, at the end, the full proposal state
    TTree tree("SimpleMCMC", "Tree of accepted points");
    sMCMC::Vector point(7, 0.1);
    {
        sMCMC::TSimpleMCMC<TUnitGaussLogLikelihood> mcmc(&tree);
        mcmc.SetChains(chains);
        mcmc.SetSeed(31);
        mcmc.GetLogLikelihood().SetDim(7);
        mcmc.Start(point, false);
        for (int i = 0; i < steps; ++i) mcmc.Step();
        mcmc.SaveStep();
    }
    // second run: a new sampler continues from the tree
    sMCMC::TSimpleMCMC<TUnitGaussLogLikelihood> again(&tree);
    again.SetChains(chains);
    again.SetSeed(31);
    again.GetLogLikelihood().SetDim(7);
    sMCMC::Vector zero(7, 0.0);
    again.Start(zero, false);
    again.Restore(&tree);
    std::printf("restored llh %.17g saved %.17g x0 %.17g sigma %.17g trials %d\n", again.GetAcceptedLogLikelihood(chains - 1),
                tree.DoubleColumn("LogLikelihood").back(), again.GetAccepted(chains - 1)[0],
                again.GetProposeStep().GetSigma(chains - 1), again.GetProposeStep().GetTrials(chains - 1));
    for (int i = 0; i < 50; ++i) {
        again.Step();
        std::printf("step %d acc %d llh %.17g\n", i, (int)again.GetStepAccepted()[chains - 1],
                    again.GetAcceptedLogLikelihood(chains - 1));
    }
    return 0;
}

int main(int argc, char** argv) {
    const char* kind = argc > 1 ? argv[1] : "hmc";
    int chains = argc > 2 ? std::atoi(argv[2]) : 1;
    int steps = argc > 3 ? std::atoi(argv[3]) : 100;
    if (!std::strcmp(kind, "restore")) return RunRestore(chains, steps);
    return RunHmc(chains, steps);
}
