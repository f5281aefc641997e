// rootshim TRandom: same call surface as ROOT's TRandom (Rndm, Uniform, Gaus,
// Exp, SetSeed) with a virtual Rndm()/Gaus() so that tests can install an
// injected, counter-addressed stream (oracle/ref_driver.cc) through gRandom.
// ROOT behaviour restated [from memory, see SURVEY.md A.6]:
//   Uniform(x1) = x1*Rndm(); Uniform(x1,x2) = x1+(x2-x1)*Rndm();
//   Exp(tau) = -tau*log(Rndm()); Gaus(mean,sigma) = mean + sigma*g.
#ifndef ROOTSHIM_TRandom_h
#define ROOTSHIM_TRandom_h
#include <cmath>
#include <cstdint>
#include <random>
#include "TObject.h"
#include "smcmc_rng.h"

class TRandom : public TObject {
public:
    TRandom(unsigned long seed = 65539) { SetSeed(seed); }
    virtual ~TRandom() {}
    virtual void SetSeed(unsigned long seed = 0) {
        if (seed == 0) { std::random_device rd; seed = rd(); }
        fEngine.seed(seed);
    }
    // Uniform deviate in the open interval (0,1), 52 bits.
    virtual double Rndm() {
        uint64_t k = fEngine();
        return smcmc_bits_to_open01((uint32_t)(k >> 32), (uint32_t)k);
    }
    // Unit normal deviate.  The default generator uses the same
    // deterministic Box-Muller kernel as the device stream.
    virtual double UnitGaus() {
        uint64_t a = fEngine(), b = fEngine();
        smcmc_u32x4 bits;
        bits.v[0] = (uint32_t)(a >> 32); bits.v[1] = (uint32_t)a;
        bits.v[2] = (uint32_t)(b >> 32); bits.v[3] = (uint32_t)b;
        double z0 = 0.0, z1 = 0.0;
        smcmc_normal_pair_from_bits(bits, 1, &z0, &z1);
        return z0;
    }
    double Uniform(double x1 = 1.0) { return x1 * Rndm(); }
    double Uniform(double x1, double x2) { return x1 + (x2 - x1) * Rndm(); }
    double Gaus(double mean = 0.0, double sigma = 1.0) {
        return mean + sigma * UnitGaus();
    }
    double Exp(double tau) { return -tau * std::log(Rndm()); }
private:
    std::mt19937_64 fEngine;
};

inline TRandom* gRandom = new TRandom(65539);
#endif
