// builtin_exprs_f.cu -- the point-wise family on its own (one operand per node), pre-instantiated for 2-D fields so that the ctypes
// tests can exercise every node of opf_device.cuh against the oracle (tests/test_gpu_pointwise.py).  The C++ front-end instantiates
// whatever a user program combines them into.
#include "engine.hpp"

namespace opfe {
    void register_builtin(const char* sig, opf_expr_launcher fn);
}
using namespace opf;

#define OPF_CAT2(a, b) a##b
#define OPF_CAT(a, b) OPF_CAT2(a, b)
#define OPF_BUILTIN_2D(...)                                                                                            \
    static const int OPF_CAT(opf_reg_, __COUNTER__) = (opfe::register_builtin(#__VA_ARGS__, &opf::launcher<__VA_ARGS__, 2>), 0);

OPF_BUILTIN_2D(Exp<F<0>>)
OPF_BUILTIN_2D(Log<F<0>>)
OPF_BUILTIN_2D(Sin<F<0>>)
OPF_BUILTIN_2D(Cos<F<0>>)
OPF_BUILTIN_2D(Tan<F<0>>)
OPF_BUILTIN_2D(Tanh<F<0>>)
OPF_BUILTIN_2D(Not<F<0>>)
OPF_BUILTIN_2D(Pos<F<0>>)
OPF_BUILTIN_2D(Exp2<F<0>>)
OPF_BUILTIN_2D(Expm1<F<0>>)
OPF_BUILTIN_2D(Log10<F<0>>)
OPF_BUILTIN_2D(Log2<F<0>>)
OPF_BUILTIN_2D(Log1p<F<0>>)
OPF_BUILTIN_2D(Cbrt<F<0>>)
OPF_BUILTIN_2D(ASin<F<0>>)
OPF_BUILTIN_2D(ACos<F<0>>)
OPF_BUILTIN_2D(ATan<F<0>>)
OPF_BUILTIN_2D(Sinh<F<0>>)
OPF_BUILTIN_2D(Cosh<F<0>>)
OPF_BUILTIN_2D(ASinh<F<0>>)
OPF_BUILTIN_2D(ACosh<F<0>>)
OPF_BUILTIN_2D(ATanh<F<0>>)
OPF_BUILTIN_2D(Erf<F<0>>)
OPF_BUILTIN_2D(Erfc<F<0>>)
OPF_BUILTIN_2D(TGamma<F<0>>)
OPF_BUILTIN_2D(LGamma<F<0>>)
OPF_BUILTIN_2D(Ceil<F<0>>)
OPF_BUILTIN_2D(Floor<F<0>>)
OPF_BUILTIN_2D(Trunc<F<0>>)
OPF_BUILTIN_2D(Round<F<0>>)
OPF_BUILTIN_2D(LRound<F<0>>)
OPF_BUILTIN_2D(LLRound<F<0>>)
OPF_BUILTIN_2D(NearbyInt<F<0>>)
OPF_BUILTIN_2D(Rint<F<0>>)
OPF_BUILTIN_2D(LRint<F<0>>)
OPF_BUILTIN_2D(LLRint<F<0>>)
OPF_BUILTIN_2D(ILogb<F<0>>)
OPF_BUILTIN_2D(Logb<F<0>>)
OPF_BUILTIN_2D(Pow<F<0>, F<1>>)
OPF_BUILTIN_2D(Min<F<0>, F<1>>)
OPF_BUILTIN_2D(FMod<F<0>, F<1>>)
OPF_BUILTIN_2D(Remainder<F<0>, F<1>>)
OPF_BUILTIN_2D(FDim<F<0>, F<1>>)
OPF_BUILTIN_2D(Hypot<F<0>, F<1>>)
OPF_BUILTIN_2D(ATan2<F<0>, F<1>>)
OPF_BUILTIN_2D(Ldexp<F<0>, F<1>>)
OPF_BUILTIN_2D(Scalbn<F<0>, F<1>>)
OPF_BUILTIN_2D(Scalbln<F<0>, F<1>>)
OPF_BUILTIN_2D(Nextafter<F<0>, F<1>>)
OPF_BUILTIN_2D(Nexttoward<F<0>, F<1>>)
OPF_BUILTIN_2D(Copysing<F<0>, F<1>>)
OPF_BUILTIN_2D(Lt<F<0>, F<1>>)
OPF_BUILTIN_2D(Le<F<0>, F<1>>)
OPF_BUILTIN_2D(Ge<F<0>, F<1>>)
OPF_BUILTIN_2D(Eq<F<0>, F<1>>)
OPF_BUILTIN_2D(Ne<F<0>, F<1>>)
OPF_BUILTIN_2D(And<F<0>, F<1>>)
OPF_BUILTIN_2D(Or<F<0>, F<1>>)
