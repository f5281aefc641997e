#ifndef ROOTSHIM_TRandom3_h
#define ROOTSHIM_TRandom3_h
#include "TRandom.h"
// Seed 0 means "pick a unique seed" as in ROOT (SimpleMCMC.C:69 relies on it).
class TRandom3 : public TRandom {
public:
    TRandom3(unsigned long seed = 4357) : TRandom(seed) {}
};
#endif
