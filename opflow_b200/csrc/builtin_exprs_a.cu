// builtin_exprs_a.cu -- device kernels for the expressions of the acceptance programs, instantiated by nvcc from the
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


// ---- solver vector kernels (engine_solver.cu): weighted-Jacobi update, diagonal scaling
OPF_BUILTIN(Mul<S<0>, Mul<F<0>, F<1>>>)
OPF_BUILTIN(Add<F<0>, Mul<S<0>, Mul<F<1>, F<2>>>>)
