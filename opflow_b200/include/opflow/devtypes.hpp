// opflow/devtypes.hpp -- the device functor node templates the front-end maps expression types onto.
// Under nvcc the real definitions (functors + kernel skeletons + launcher) come from the engine's opf_device.cuh, and every
// expression type the program assigns gets its kernels instantiated in the user's translation unit and registered with
// opf_expr_register.  Under a host-only compiler only the names are declared: the program can then use the expressions
// libopflow_b200.so carries (opf_expr_builtin_name), anything else fails loudly at run time (OPF_ERR_UNSUPPORTED).
#pragma once
#if defined(__CUDACC__) && !defined(OPFLOW_NO_KERNEL_INSTANTIATION)
#include <opf_device.cuh>
#define OPFLOW_DEVICE_KERNELS 1
#else
namespace opf {
    template <int K> struct F;
    template <int K> struct S;
    template <class L, class R> struct Add;
    template <class L, class R> struct Sub;
    template <class L, class R> struct Mul;
    template <class L, class R> struct Div;
    template <class L, class R> struct Min;
    template <class L, class R> struct Max;
    template <class L, class R> struct Pow;
    template <class L, class R> struct Lt;
    template <class L, class R> struct Le;
    template <class L, class R> struct Gt;
    template <class L, class R> struct Ge;
    template <class L, class R> struct Eq;
    template <class L, class R> struct Ne;
    template <class L, class R> struct And;
    template <class L, class R> struct Or;
    template <class E> struct Neg;
    template <class E> struct Pos;
    template <class E> struct Not;
    template <class E> struct Sqrt;
    template <class E> struct Abs;
    template <class E> struct Exp;
    template <class E> struct Log;
    template <class E> struct Sin;
    template <class E> struct Cos;
    template <class E> struct Tan;
    template <class E> struct Tanh;
    template <class E> struct Pow2;
    template <class E> struct Exp2;
    template <class E> struct Expm1;
    template <class E> struct Log10;
    template <class E> struct Log2;
    template <class E> struct Log1p;
    template <class E> struct Cbrt;
    template <class E> struct ASin;
    template <class E> struct ACos;
    template <class E> struct ATan;
    template <class E> struct Sinh;
    template <class E> struct Cosh;
    template <class E> struct ASinh;
    template <class E> struct ACosh;
    template <class E> struct ATanh;
    template <class E> struct Erf;
    template <class E> struct Erfc;
    template <class E> struct TGamma;
    template <class E> struct LGamma;
    template <class E> struct Ceil;
    template <class E> struct Floor;
    template <class E> struct Trunc;
    template <class E> struct Round;
    template <class E> struct LRound;
    template <class E> struct LLRound;
    template <class E> struct NearbyInt;
    template <class E> struct Rint;
    template <class E> struct LRint;
    template <class E> struct LLRint;
    template <class E> struct ILogb;
    template <class E> struct Logb;
    template <class L, class R> struct FMod;
    template <class L, class R> struct Remainder;
    template <class L, class R> struct FDim;
    template <class L, class R> struct Hypot;
    template <class L, class R> struct ATan2;
    template <class L, class R> struct Ldexp;
    template <class L, class R> struct Scalbn;
    template <class L, class R> struct Scalbln;
    template <class L, class R> struct Nextafter;
    template <class L, class R> struct Nexttoward;
    template <class L, class R> struct Copysing;
    template <class C, class A, class B> struct Cond;
    template <int D, class E> struct D2C;
    template <int D, class E> struct D1C;
    template <int D, class E> struct D1Dn;
    template <int D, class E> struct D1Up;
    template <int D, class E> struct WenoDn;
    template <int D, class E> struct WenoUp;
    template <int D, class E> struct IntpC2N;
    template <int D, class E> struct IntpN2C;
    template <int D, class U, class E> struct FlCentralC2N;
    template <int D, class U, class E> struct FlCentralN2C;
    template <int D, class U, class E> struct FlQuickC2N;
    template <int D, class U, class E> struct FlQuickN2C;
    template <int D, class U, class E> struct FlCuiC2N;
    template <int D, class U, class E> struct FlCuiN2C;
    template <int D, class U, class E> struct FlFrommC2N;
    template <int D, class U, class E> struct FlFrommN2C;
    template <int D, class U, class E> struct FlLuiC2N;
    template <int D, class U, class E> struct FlLuiN2C;
    template <int D, class U, class E> struct FlMinmodC2N;
    template <int D, class U, class E> struct FlMinmodN2C;
    template <int D, class U, class E> struct FlSuperbeeC2N;
    template <int D, class U, class E> struct FlSuperbeeN2C;
    template <int D, class U, class E> struct FlMusclC2N;
    template <int D, class U, class E> struct FlMusclN2C;
    template <int D, class U, class E> struct FlHarmonicC2N;
    template <int D, class U, class E> struct FlHarmonicN2C;
    template <int D, class U, class E> struct FlAlbadaC2N;
    template <int D, class U, class E> struct FlAlbadaN2C;
    template <int N0, int N1, int N2, int K0, class E> struct Conv;
    template <class Fn, class E> struct Adapt1;
    template <class Fn, class L, class R> struct Adapt2;
}// namespace opf
#endif
