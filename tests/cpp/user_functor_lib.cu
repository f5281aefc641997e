// user_functor_lib.cu -- a user translation unit compiled by nvcc into
// libuser_functor.so: it instantiates the kernels of smcmc_device_functor.cuh for
// the user's functor and registers their launch table with an engine created by
// anyone (here: the Python tests, through ctypes).
#include "constrained_functor.cuh"

static smcmc_user::Binding<TConstrainedLikelihood>* gBinding = nullptr;

extern "C" int user_constrained_bind(smcmc_engine* e) {
    try {
        TConstrainedLikelihood like;
        like.Init();
        if (!gBinding) gBinding = new smcmc_user::Binding<TConstrainedLikelihood>();
        gBinding->Bind(e, like);
        return 0;
    } catch (std::exception&) {
        return -1;
    }
}
