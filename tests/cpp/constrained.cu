// constrained.cu -- the reference's example4/Constrained.C written against the
// mirror header with the likelihood as a USER device functor
// (constrained_functor.cuh): the template contract
//     sMCMC::TSimpleMCMC<TConstrainedLikelihood> mcmc(tree);
// is the reference's (example4/Constrained.C:17), the Step hot path runs on the GPU.
//   constrained chain <steps>             one chain = chain 3 of seed 51 (golden "constrained25")
//   constrained ensemble <chains> <steps> posterior mean / covariance of the ensemble
//   constrained hmc <chains> <steps>      TSimpleHMC<L, L> with the functor's own gradient
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "constrained_functor.cuh"
#include "TSimpleHMC.H"

static int Chain(int steps) {
    TTree tree("Constrained", "Tree of accepted points");
    sMCMC::TSimpleMCMC<TConstrainedLikelihood> mcmc(&tree);
    mcmc.SetSeed(51);
    mcmc.SetDevice(0, 3);
    TConstrainedLikelihood& like = mcmc.GetLogLikelihood();
    like.Init();
    mcmc.GetProposeStep().SetDim(like.GetDim());
    sMCMC::Vector p(like.GetDim(), 70.0);
    if (!mcmc.Start(p, false)) return 2;
    std::printf("start llh %.17g direct %.17g\n", mcmc.GetAcceptedLogLikelihood(), like(p));
    for (int i = 0; i < steps; ++i) {
        const bool ok = mcmc.Step();
        std::printf("step %d acc %d llh %.17g x0 %.17g sigma %.17g\n", i, (int)ok, mcmc.GetAcceptedLogLikelihood(),
                    mcmc.GetAccepted()[0], mcmc.GetProposeStep().GetSigma());
    }
    std::printf("entries %ld calls %d\n", tree.GetEntries(), mcmc.GetLogLikelihoodCount());
    return 0;
}

// The ensemble version of example4: every chain runs the schedule of Constrained.C
// (burn-in, UpdateProposal, burn-in, UpdateProposal, run) and the accepted points of
// the last `steps` steps are averaged over chains and steps, which is what
// ConstrainedCheck.C:19-69 profiles from the tree.
static int Ensemble(int chains, int steps) {
    sMCMC::TSimpleMCMC<TConstrainedLikelihood> mcmc;
    mcmc.SetChains(chains);
    mcmc.SetSeed(52);
    TConstrainedLikelihood& like = mcmc.GetLogLikelihood();
    like.Init();
    const int n = (int)like.GetDim();
    mcmc.GetProposeStep().SetDim(n);
    sMCMC::Vector p(n, 76.0);
    if (!mcmc.Start(p, false)) return 2;
    mcmc.StepMany(4000);
    mcmc.GetProposeStep().UpdateProposal();
    mcmc.StepMany(6000);
    mcmc.GetProposeStep().UpdateProposal();
    std::vector<double> mean(n, 0.0), second((size_t)n * n, 0.0);
    double sumMean = 0.0, sumSq = 0.0;
    long count = 0;
    for (int s = 0; s < steps; ++s) {
        mcmc.StepMany(25);
        for (int c = 0; c < chains; ++c) {
            const sMCMC::Vector x = mcmc.GetAccepted(c);
            double sum = 0.0;
            for (int i = 0; i < n; ++i) {
                mean[i] += x[i];
                sum += x[i];
                for (int j = 0; j < n; ++j) second[(size_t)i * n + j] += x[i] * x[j];
            }
            sumMean += sum;
            sumSq += sum * sum;
            ++count;
        }
    }
    std::printf("samples %ld\n", count);
    for (int i = 0; i < n; ++i) std::printf("mean %d %.10g var %.10g\n", i, mean[i] / count,
                                            second[(size_t)i * n + i] / count - mean[i] / count * mean[i] / count);
    std::printf("cov01 %.10g\n", second[1] / count - mean[0] / count * mean[1] / count);
    std::printf("sum mean %.10g var %.10g\n", sumMean / count, sumSq / count - sumMean / count * sumMean / count);
    return 0;
}

static int Hmc(int chains, int steps) {
    sMCMC::TSimpleHMC<TConstrainedLikelihood, TConstrainedLikelihood> hmc;
    hmc.SetChains(chains);
    hmc.SetSeed(53);
    TConstrainedLikelihood& like = hmc.GetLogLikelihood();
    like.Init();
    const int n = (int)like.GetDim();
    sMCMC::Vector p(n, 76.0);
    hmc.Start(p, false);
    for (int i = 0; i < steps; ++i) hmc.Step(false);
    std::printf("hmc potentials %d gradients %d acceptance %.6g\n", hmc.GetPotentialCount(), hmc.GetGradientCount(),
                hmc.GetAcceptanceRate());
    const sMCMC::Vector x = hmc.GetAccepted();
    double sum = 0.0;
    for (int i = 0; i < n; ++i) sum += x[i];
    std::printf("hmc sum %.10g x24 %.10g\n", sum, x[24]);
    return 0;
}

int main(int argc, char** argv) {
    const char* mode = argc > 1 ? argv[1] : "chain";
    try {
        if (!std::strcmp(mode, "ensemble")) return Ensemble(argc > 2 ? std::atoi(argv[2]) : 256, argc > 3 ? std::atoi(argv[3]) : 40);
        if (!std::strcmp(mode, "hmc")) return Hmc(argc > 2 ? std::atoi(argv[2]) : 4, argc > 3 ? std::atoi(argv[3]) : 50);
        return Chain(argc > 2 ? std::atoi(argv[2]) : 100);
    } catch (std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
