// builtin_exprs_d.cu -- device kernels for the expressions of the acceptance programs, instantiated by nvcc from the
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

// ---- examples/LidDriven/LidDriven2D.cpp:83-90 explicit updates (BASELINE config C4)
OPF_BUILTIN(Sub<F<0>, Mul<S<0>, D1C<1, Mul<IntpC2N<1, F<1>>, IntpC2N<0, F<2>>>>>>)// du - 0.5*dt*conv_xy(u,dv)
OPF_BUILTIN(Sub<F<0>, Mul<S<0>, D1C<0, F<1>>>>)                                   // u - dt*dx(dp)
OPF_BUILTIN(Sub<F<0>, Mul<S<0>, D1C<1, F<1>>>>)                                   // v - dt*dy(dp)
OPF_BUILTIN(Sub<F<0>, Mul<S<0>, D1C<2, F<1>>>>)                                   // w - dt*dz(dp)
OPF_BUILTIN(Div<Add<D1C<0, F<0>>, D1C<1, F<1>>>, S<0>>)                           // Poisson rhs (dx(du)+dy(dv))/dt
OPF_BUILTIN(Div<Add<Add<D1C<0, F<0>>, D1C<1, F<1>>>, D1C<2, F<2>>>, S<0>>)
OPF_BUILTIN(D1C<0, Mul<IntpN2C<0, F<0>>, IntpN2C<0, F<1>>>>)                      // conv_xx
OPF_BUILTIN(D1C<1, Mul<IntpC2N<1, F<0>>, IntpC2N<0, F<1>>>>)                      // conv_xy
OPF_BUILTIN(D1C<0, Mul<IntpC2N<1, F<0>>, IntpC2N<0, F<1>>>>)                      // conv_yx
OPF_BUILTIN(D1C<1, Mul<IntpN2C<1, F<0>>, IntpN2C<1, F<1>>>>)                      // conv_yy

