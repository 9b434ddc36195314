// builtin_exprs.cu -- device kernels for the expressions of the acceptance programs, instantiated by nvcc from the
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

// ---- trivial assignments: field = c (CartesianField.hpp:237), field = field (:180), compound forms
OPF_BUILTIN(S<0>)
OPF_BUILTIN(F<0>)
OPF_BUILTIN(Add<F<0>, F<1>>)
OPF_BUILTIN(Sub<F<0>, F<1>>)
OPF_BUILTIN(Mul<F<0>, F<1>>)
OPF_BUILTIN(Div<F<0>, F<1>>)
OPF_BUILTIN(Mul<S<0>, F<0>>)
OPF_BUILTIN(Add<F<0>, S<0>>)
OPF_BUILTIN(Sub<F<0>, S<0>>)
OPF_BUILTIN(Add<F<0>, Mul<S<0>, F<1>>>)
OPF_BUILTIN(Pow2<F<0>>)
OPF_BUILTIN(Abs<F<0>>)
OPF_BUILTIN(Sqrt<F<0>>)
OPF_BUILTIN(Neg<F<0>>)
OPF_BUILTIN(Abs<Sub<F<0>, F<1>>>)
OPF_BUILTIN(Cond<Gt<F<0>, S<0>>, F<1>, F<2>>)
OPF_BUILTIN(Max<F<0>, F<1>>)
OPF_BUILTIN(Min<F<0>, S<0>>)

// ---- single operators (tests/: one kernel per reference Op::eval)
OPF_BUILTIN(D2C<0, F<0>>)
OPF_BUILTIN(D2C<1, F<0>>)
OPF_BUILTIN(D2C<2, F<0>>)
OPF_BUILTIN(D1C<0, F<0>>)
OPF_BUILTIN(D1C<1, F<0>>)
OPF_BUILTIN(D1C<2, F<0>>)
OPF_BUILTIN(D1Dn<0, F<0>>)
OPF_BUILTIN(D1Dn<1, F<0>>)
OPF_BUILTIN(D1Up<0, F<0>>)
OPF_BUILTIN(D1Up<1, F<0>>)
OPF_BUILTIN(WenoDn<0, F<0>>)
OPF_BUILTIN(WenoUp<0, F<0>>)
OPF_BUILTIN(WenoDn<1, F<0>>)
OPF_BUILTIN(WenoUp<1, F<0>>)
OPF_BUILTIN(IntpC2N<0, F<0>>)
OPF_BUILTIN(IntpC2N<1, F<0>>)
OPF_BUILTIN(IntpN2C<0, F<0>>)
OPF_BUILTIN(IntpN2C<1, F<0>>)

// ---- examples/FTCS/FTCS.cpp, examples/FTCS2D/FTCS-OMP.cpp:26 and its 3-D extension (BASELINE configs C1, C2)
OPF_BUILTIN(Add<F<0>, Mul<S<0>, D2C<0, F<1>>>>)
OPF_BUILTIN(Add<F<0>, Mul<S<0>, Add<D2C<0, F<1>>, D2C<1, F<2>>>>>)
OPF_BUILTIN(Add<F<0>, Mul<S<0>, Add<Add<D2C<0, F<1>>, D2C<1, F<2>>>, D2C<2, F<3>>>>>)
// Laplacian (benchmark/Core/LaplaceOp.cpp:80 shape) and the Poisson operator of LidDriven (LidDriven2D.cpp:70)
OPF_BUILTIN(Add<D2C<0, F<0>>, D2C<1, F<1>>>)
OPF_BUILTIN(Add<Add<D2C<0, F<0>>, D2C<1, F<1>>>, D2C<2, F<2>>>)

// ---- examples/CONV1D/CONV1D.cpp:29-31 (BASELINE config C3)
OPF_BUILTIN(Sub<F<0>, Mul<S<0>, D1Dn<0, F<1>>>>)
OPF_BUILTIN(Sub<F<0>, Mul<S<0>, WenoDn<0, F<1>>>>)
OPF_BUILTIN(Add<F<0>, Mul<S<0>, WenoUp<0, F<1>>>>)
OPF_BUILTIN(Sub<F<0>, Mul<S<0>, WenoDn<1, F<1>>>>)

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
