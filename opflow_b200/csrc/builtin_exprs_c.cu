// builtin_exprs_c.cu -- device kernels for the expressions of the acceptance programs, instantiated by nvcc from the
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

// ---- examples/FTCS/FTCS.cpp, examples/FTCS2D/FTCS-OMP.cpp:26 and its 3-D extension (BASELINE configs C1, C2)
OPF_BUILTIN(Add<F<0>, Mul<S<0>, D2C<0, F<1>>>>)
OPF_BUILTIN(Add<F<0>, Mul<S<0>, Add<D2C<0, F<1>>, D2C<1, F<2>>>>>)
// (the 3-D form, BASELINE config C2, lives in builtin_exprs_e.cu)
// Laplacian (benchmark/Core/LaplaceOp.cpp:80 shape) and the Poisson operator of LidDriven (LidDriven2D.cpp:70)
OPF_BUILTIN(Add<D2C<0, F<0>>, D2C<1, F<1>>>)
OPF_BUILTIN(Add<Add<D2C<0, F<0>>, D2C<1, F<1>>>, D2C<2, F<2>>>)
// fused residual r = b - L(x) of the Poisson operator (engine_solver.cu)
OPF_BUILTIN(Sub<F<0>, Add<D2C<0, F<1>>, D2C<1, F<2>>>>)
OPF_BUILTIN(Sub<F<0>, Add<Add<D2C<0, F<1>>, D2C<1, F<2>>>, D2C<2, F<3>>>>)
// fused weighted-Jacobi sweep x <- x + w*dinv*(b - L(x)) of the Poisson operator (engine_solver.cu smooth())
OPF_BUILTIN(Add<F<0>, Mul<S<0>, Mul<F<1>, Sub<F<2>, Add<D2C<0, F<0>>, D2C<1, F<0>>>>>>>)
OPF_BUILTIN(Add<F<0>, Mul<S<0>, Mul<F<1>, Sub<F<2>, Add<Add<D2C<0, F<0>>, D2C<1, F<0>>>, D2C<2, F<0>>>>>>>)
// red-black Gauss-Seidel half-sweeps (PFMG relax_type 2 / 3): x = Par.dinv.b from a zero guess, x += Par.dinv.r from a residual, and
// the fused form  x += Par.dinv.(b - L(x))  of the Poisson operator (engine_solver.cu smooth())
OPF_BUILTIN(Mul<Par<0>, Mul<F<0>, F<1>>>)
OPF_BUILTIN(Add<F<0>, Mul<Par<0>, Mul<F<1>, F<2>>>>)
OPF_BUILTIN(Add<F<0>, Mul<Par<1>, Mul<F<1>, F<2>>>>)
OPF_BUILTIN(Add<F<0>, Mul<Par<0>, Mul<F<1>, Sub<F<2>, Add<D2C<0, F<0>>, D2C<1, F<0>>>>>>>)
OPF_BUILTIN(Add<F<0>, Mul<Par<1>, Mul<F<1>, Sub<F<2>, Add<D2C<0, F<0>>, D2C<1, F<0>>>>>>>)
OPF_BUILTIN(Add<F<0>, Mul<Par<0>, Mul<F<1>, Sub<F<2>, Add<Add<D2C<0, F<0>>, D2C<1, F<0>>>, D2C<2, F<0>>>>>>>)
OPF_BUILTIN(Add<F<0>, Mul<Par<1>, Mul<F<1>, Sub<F<2>, Add<Add<D2C<0, F<0>>, D2C<1, F<0>>>, D2C<2, F<0>>>>>>>)
// the equation of the reference's CSR generator test in residual form, `1.0 == d2x(e) + d2y(e)` -> 1 - L(e) (CSRMatrixGeneratorTest.cpp:57-158)
OPF_BUILTIN(Sub<S<0>, Add<D2C<0, F<0>>, D2C<1, F<1>>>>)
// 1-D Poisson / Helmholtz pieces
OPF_BUILTIN(Sub<F<0>, D2C<0, F<1>>>)

// ---- examples/CONV1D/CONV1D.cpp:29-31 (BASELINE config C3)
OPF_BUILTIN(Sub<F<0>, Mul<S<0>, D1Dn<0, F<1>>>>)
OPF_BUILTIN(Sub<F<0>, Mul<S<0>, WenoDn<0, F<1>>>>)
OPF_BUILTIN(Add<F<0>, Mul<S<0>, WenoUp<0, F<1>>>>)
OPF_BUILTIN(Sub<F<0>, Mul<S<0>, WenoDn<1, F<1>>>>)

