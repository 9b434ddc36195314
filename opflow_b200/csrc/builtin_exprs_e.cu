// builtin_exprs_e.cu -- device kernels for the expressions of the acceptance programs, instantiated by nvcc from the
// functor templates in opf_device.cuh and registered under their signature at load time.  User programs compiled with
// nvcc against <OpFlow> register their own expression types the same way (opf_expr_register).
#include "engine.hpp"

namespace opfe {
    void register_builtin(const char* sig, opf_expr_launcher fn);
}
using namespace opf;

#define OPF_CAT2(a, b) a##b
#define OPF_CAT(a, b) OPF_CAT2(a, b)
#define OPF_BUILTIN(...)                                                                                               \
    static const int OPF_CAT(opf_reg_, __COUNTER__) = (opfe::register_builtin(#__VA_ARGS__, &opf::launcher<__VA_ARGS__>), 0);

// ---- BASELINE config C2: 3-D FTCS heat equation, the 3-D extension of examples/FTCS2D/FTCS-OMP.cpp:26 (the benchmark kernel)
OPF_BUILTIN(Add<F<0>, Mul<S<0>, Add<Add<D2C<0, F<1>>, D2C<1, F<2>>>, D2C<2, F<3>>>>>)
