// builtin_exprs_b.cu -- device kernels for the expressions of the acceptance programs, instantiated by nvcc from the
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

