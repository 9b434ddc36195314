// builtin_exprs_g.cu -- flux-limiter interpolators and convolutions on their own, pre-instantiated for the ctypes tests
// (tests/test_gpu_limiters.py): every scheme of D1FluxLimiterBasedIntpOp.hpp:22-61 along x in both directions (2-D fields), two
// schemes along y, and the 3 x 3 / 3 x 3 x 3 convolutions of examples/LevelSet/UniLS.cpp:107-113.
#include "engine.hpp"

namespace opfe {
    void register_builtin(const char* sig, opf_expr_launcher fn);
}
using namespace opf;

#define OPF_CAT2(a, b) a##b
#define OPF_CAT(a, b) OPF_CAT2(a, b)
#define OPF_BUILTIN_2D(...)                                                                                            \
    static const int OPF_CAT(opf_reg_, __COUNTER__) = (opfe::register_builtin(#__VA_ARGS__, &opf::launcher<__VA_ARGS__, 2>), 0);
#define OPF_BUILTIN_3D(...)                                                                                            \
    static const int OPF_CAT(opf_reg_, __COUNTER__) = (opfe::register_builtin(#__VA_ARGS__, &opf::launcher<__VA_ARGS__, 4>), 0);

OPF_BUILTIN_2D(FlCentralC2N<0, F<0>, F<1>>)
OPF_BUILTIN_2D(FlCentralN2C<0, F<0>, F<1>>)
OPF_BUILTIN_2D(FlQuickC2N<0, F<0>, F<1>>)
OPF_BUILTIN_2D(FlQuickN2C<0, F<0>, F<1>>)
OPF_BUILTIN_2D(FlCuiC2N<0, F<0>, F<1>>)
OPF_BUILTIN_2D(FlCuiN2C<0, F<0>, F<1>>)
OPF_BUILTIN_2D(FlFrommC2N<0, F<0>, F<1>>)
OPF_BUILTIN_2D(FlFrommN2C<0, F<0>, F<1>>)
OPF_BUILTIN_2D(FlLuiC2N<0, F<0>, F<1>>)
OPF_BUILTIN_2D(FlLuiN2C<0, F<0>, F<1>>)
OPF_BUILTIN_2D(FlMinmodC2N<0, F<0>, F<1>>)
OPF_BUILTIN_2D(FlMinmodN2C<0, F<0>, F<1>>)
OPF_BUILTIN_2D(FlSuperbeeC2N<0, F<0>, F<1>>)
OPF_BUILTIN_2D(FlSuperbeeN2C<0, F<0>, F<1>>)
OPF_BUILTIN_2D(FlMusclC2N<0, F<0>, F<1>>)
OPF_BUILTIN_2D(FlMusclN2C<0, F<0>, F<1>>)
OPF_BUILTIN_2D(FlHarmonicC2N<0, F<0>, F<1>>)
OPF_BUILTIN_2D(FlHarmonicN2C<0, F<0>, F<1>>)
OPF_BUILTIN_2D(FlAlbadaC2N<0, F<0>, F<1>>)
OPF_BUILTIN_2D(FlAlbadaN2C<0, F<0>, F<1>>)
OPF_BUILTIN_2D(FlQuickC2N<1, F<0>, F<1>>)
OPF_BUILTIN_2D(FlMinmodN2C<1, F<0>, F<1>>)
OPF_BUILTIN_3D(FlSuperbeeC2N<2, F<0>, F<1>>)
OPF_BUILTIN_2D(Conv<3, 3, 1, 0, F<0>>)
OPF_BUILTIN_2D(Conv<5, 3, 1, 0, F<0>>)
OPF_BUILTIN_3D(Conv<3, 3, 3, 0, F<0>>)
OPF_BUILTIN_2D(Mul<S<0>, Conv<3, 3, 1, 1, Mul<F<0>, F<1>>>>)
