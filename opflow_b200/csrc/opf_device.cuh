// opf_device.cuh -- device functors + kernel skeletons of the B200 evaluation engine (sm_100a).
//
// An OpFlow expression tree (reference: Expression<Op, Args...>, src/Core/Expr/Expression.hpp:24-111) becomes a
// *stateless type* built from the node templates below; all run-time data (field pointers, scalars, per-axis mesh
// arrays, LocOnMesh bits) travels in one POD `ExprArgs` passed by value as a __grid_constant__ kernel parameter, so every
// leaf access is a constant-bank operand and every node index is a compile-time constant.
//
// Arithmetic policies (include/opflow_b200.h opf_mode):
//   Exact -- IEEE-rn add/sub/mul/div intrinsics in exactly the reference's operation order (no FMA contraction):
//            bit-identical to the reference's g++ -O3 x86-64 CPU path.
//   Fast  -- divides by mesh spacings become multiplies by per-axis reciprocal arrays, FMA contraction allowed
//            (<= 1e-12 relative vs the reference).  On single-spacing axes (AxisView::uniform) the coefficients are kernel-
//            parameter constants: no coefficient loads at all (the UNI variants of the window and TMA skeletons).
//
// Skeletons (K1 of SURVEY 2.3): assign_kernel (one thread per cell; 1-D and small / unaligned boxes), window_kernel
// (per-thread register window marching the slow axis; 2-D), tma_kernel (TMA-streamed tile-plus-halo planes in a shared-memory
// ring, own cells by LDS.128, z-planes rotated in registers; 3-D).  launcher<E, DIMS> dispatches between them.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <cstdlib>

// run-time switches and launch bookkeeping live in libopflow_b200.so (engine_core.cu); launchers instantiated in a user's
// translation unit reach them through these two C symbols
enum { OPF_OPT_TMA = 0, OPF_OPT_TMA2D, OPF_OPT_WINDOW, OPF_OPT_OVERLAP, OPF_OPT_GRAPHS, OPF_OPT_MG_FUSED, OPF_OPT_DIRECT_HALO,
       OPF_OPT_FUSED_KRYLOV, OPF_OPT_COUNT };
extern "C" int opf_internal_opt(int id);
extern "C" void opf_internal_note_kernel(const char* name);

namespace opf {

    constexpr int MAX_FIELDS = 32;
    constexpr int MAX_SCALARS = 32;// a 3 x 3 x 3 convolution kernel occupies 27 slots
    constexpr int MAX_NODES = 96;

    // element (i0,i1,i2) (global indices) lives at p[i0 + i1*s1 + i2*s2]; the field offset is folded into p
    // (reference: data[i - offset], CartesianField.hpp:775-780 + PlainTensor::getOffset PlainTensor.hpp:204-212)
    struct FieldView {
        const double* p;
        long long s1, s2;
    };
    struct DstView {
        double* p;
        long long s1, s2;
    };
    // per-axis mesh arrays, each biased so that a[i] is valid for every global node index i of the mesh's ext range
    // (CartesianMesh::x/dx, CartesianMesh.hpp:44-49).  r* are the Fast-mode reciprocals, computed on the host:
    //   rdx[i]  = 1/dx[i]
    //   rdxh[i] = 1/((dx[i-1]+dx[i])*0.5)
    //   rdxc[i] = 1/((dxl+dxr)*0.5), dxl=(dx[i-1]+dx[i])*0.5, dxr=(dx[i]+dx[i+1])*0.5
    // `uniform`: every dx entry of the axis is bitwise the same number (MeshBuilder::setMeshOfDim(k, min, max),
    // CartesianMesh.hpp:292-303, and its symmetric/periodic extension); then u[CF_*] holds that one coefficient and the
    // kernels read it from the constant bank instead of loading the arrays (u[CF_RDX2] = rdx*rdx).
    struct AxisView {
        const double *x, *dx, *rdx, *rdxh, *rdxc;
        double u[6];
        int uniform;
    };
    struct ExprArgs {
        FieldView f[MAX_FIELDS];
        double s[MAX_SCALARS];
        AxisView ax[3];
        unsigned char loc[MAX_NODES];// per tree node (preorder): bit d set <=> the node's *operand* is Center on axis d
    };
    struct LaunchRange {
        int lo[3], hi[3];
    };
    // everything a launcher needs besides ExprArgs
    struct LaunchInfo {
        DstView dst;       // where the result is written
        const double* old; // current dst values for compound ops (same strides as dst); may equal dst.p
        LaunchRange r;
        int dim;
        int op;   // opf_assign_op
        int mode; // opf_mode
        int alias0;// all field leaves are the same field -> read everything through f[0]
        int uniform;    // every axis of the mesh has a single spacing (AxisView::uniform): coefficient loads become constants
        unsigned valign;// bit s: rows of field slot s are 16-byte aligned at r.lo[0] (vector loads allowed)
        int dalign;     // same for dst
        int window;     // 0: direct-global skeleton, >0: register-window skeleton allowed
        // TMA tile skeleton (3-D): per field slot the tensor behind the pitched storage; tensor coord = global index - org
        struct TmaSlot {
            const void* base;
            unsigned long long dim[3], stride_b[2];
            int org[3];
        } tma[MAX_FIELDS];
        int tma_ok;// every slot is TMA-addressable (16-byte aligned base and strides)
        // reduction launches
        int rop;
        double* partials;// >= grid blocks
        int n_partials;
        double* result;// where the folded value goes (device memory); null: partials + n_partials
    };

    // layout stamp of the blobs exchanged between libopflow_b200.so and launchers instantiated in user translation units
#define OPF_DEVICE_ABI (((unsigned long long) sizeof(opf::ExprArgs) << 32) | (unsigned long long) sizeof(opf::LaunchInfo))

    // ------------------------------------------------------------------------------------------- policies
    struct Exact {
        static constexpr bool fast = false;
        __device__ __forceinline__ static double add(double a, double b) { return __dadd_rn(a, b); }
        __device__ __forceinline__ static double sub(double a, double b) { return __dsub_rn(a, b); }
        __device__ __forceinline__ static double mul(double a, double b) { return __dmul_rn(a, b); }
        __device__ __forceinline__ static double div(double a, double b) { return __ddiv_rn(a, b); }
    };
    // Stencil -- the arithmetic of the reference's IMPLICIT path: StencilPad's `pad / num` is `pad * (1. / num)`
    // (src/DataStructures/StencilPad.hpp:293-296), so a coefficient the assembled matrix holds is a*(1/d), one rounding away from the
    // explicit path's a/d whenever d is not a power of two.  Probing the operator in this mode reproduces the reference's assembled
    // CSR / HYPRE coefficients bit for bit on any mesh (tests/test_gpu_coefficients.py against tests/golden/ref_csr.json).
    struct Stencil : Exact {
        __device__ __forceinline__ static double div(double a, double b) { return __dmul_rn(a, __ddiv_rn(1.0, b)); }
    };
    struct Fast {
        static constexpr bool fast = true;
        __device__ __forceinline__ static double add(double a, double b) { return a + b; }
        __device__ __forceinline__ static double sub(double a, double b) { return a - b; }
        __device__ __forceinline__ static double mul(double a, double b) { return a * b; }
        __device__ __forceinline__ static double div(double a, double b) { return a / b; }
    };

    template <int D>
    __device__ __forceinline__ int axis_of(int i, int j, int k) {
        return D == 0 ? i : (D == 1 ? j : k);
    }
#ifndef OPF_WIN_MINBLOCKS
#define OPF_WIN_MINBLOCKS 6
#endif
#define OPF_SHIFT(D, n) i + ((D) == 0 ? (n) : 0), j + ((D) == 1 ? (n) : 0), k + ((D) == 2 ? (n) : 0)

    template <int N>
    struct IC {
        static constexpr int value = N;
    };
    // compile-time loop: fn(IC<LO>{}) ... fn(IC<HI>{})
    template <int LO, int HI, class Fn>
    __device__ __forceinline__ void static_for(Fn&& fn) {
        if constexpr (LO <= HI) {
            fn(IC<LO>{});
            static_for<LO + 1, HI>(fn);
        }
    }

    // ------------------------------------------------------------------------------------------- tap sets
    // Compile-time footprint of an expression: for every field slot the set of relative offsets (di,dj,dk) it is read
    // at.  The register-window skeleton sizes its per-thread windows from it.  Radius is capped at WR per axis; larger
    // footprints (overflow) use the direct-global skeleton.
    constexpr int WR = 3, WN = 2 * WR + 1;
    struct TapGrid {
        bool t[WN][WN][WN] = {};// [dk+WR][dj+WR][di+WR]
        bool overflow = false;
    };
    constexpr TapGrid tap_origin() {
        TapGrid g;
        g.t[WR][WR][WR] = true;
        return g;
    }
    // Minkowski sum of a tap grid with offsets [lo,hi] along axis D
    constexpr TapGrid tap_shift(const TapGrid& in, int D, int lo, int hi) {
        TapGrid o;
        o.overflow = in.overflow;
        for (int k = 0; k < WN; ++k)
            for (int j = 0; j < WN; ++j)
                for (int i = 0; i < WN; ++i)
                    if (in.t[k][j][i])
                        for (int d = lo; d <= hi; ++d) {
                            const int ii = i + (D == 0 ? d : 0), jj = j + (D == 1 ? d : 0), kk = k + (D == 2 ? d : 0);
                            if (ii < 0 || ii >= WN || jj < 0 || jj >= WN || kk < 0 || kk >= WN) o.overflow = true;
                            else
                                o.t[kk][jj][ii] = true;
                        }
        return o;
    }
    template <int NS>
    struct TapSet {
        TapGrid g[NS > 0 ? NS : 1];
        bool overflow = false;
        constexpr void add(int slot, const TapGrid& in) {
            for (int k = 0; k < WN; ++k)
                for (int j = 0; j < WN; ++j)
                    for (int i = 0; i < WN; ++i)
                        if (in.t[k][j][i]) g[slot].t[k][j][i] = true;
            if (in.overflow) overflow = true;
        }
    };

    // ------------------------------------------------------------------------------------------- leaves
    // Every node offers two evaluators with identical arithmetic:
    //   eval<B,P,A0>(args, i,j,k)            run-time coordinates, loads straight from global memory
    //   ev<B,P,A0,DI,DJ,DK>(ctx)             compile-time offsets relative to the thread's tile origin; leaves read the
    //                                        thread's register window (ctx.get<slot,DI,DJ,DK>())
    // CartesianField::evalAtImpl_final (CartesianField.hpp:775-780)
    template <int K>
    struct F {
        static constexpr int size = 1, maxaxis = -1, nf = K + 1;
        template <int B, class P, bool A0>
        __device__ __forceinline__ static double eval(const ExprArgs& a, int i, int j, int k) {
            const FieldView& v = a.f[A0 ? 0 : K];
            return __ldg(v.p + ((long long) i + (long long) j * v.s1 + (long long) k * v.s2));
        }
        template <int B, class P, bool A0, int DI, int DJ, int DK, class C>
        __device__ __forceinline__ static double ev(const C& c) {
            return c.template get<(A0 ? 0 : K), DI, DJ, DK>();
        }
        template <bool A0, class TS>
        static constexpr void taps(TS& ts, const TapGrid& in) {
            ts.add(A0 ? 0 : K, in);
        }
    };
    // ScalarExpr<T>::evalAt ignores the index (ScalarExpr.hpp:36)
    template <int K>
    struct S {
        static constexpr int size = 1, maxaxis = -1, nf = 0;
        template <int B, class P, bool A0>
        __device__ __forceinline__ static double eval(const ExprArgs& a, int, int, int) {
            return a.s[K];
        }
        template <int B, class P, bool A0, int DI, int DJ, int DK, class C>
        __device__ __forceinline__ static double ev(const C& c) {
            return c.a.s[K];
        }
        template <bool A0, class TS>
        static constexpr void taps(TS&, const TapGrid&) {}
    };

    // Par<C>: 1.0 on the cells whose global index sum i + j + k has parity C, else 0.0 -- the colour mask of the red-black
    // Gauss-Seidel half-sweeps (HYPRE PFMG relax_type 2 / 3, StructSolverPFMG.hpp:23-34).  Carries no ranges, like a scalar.
    template <int C>
    struct Par {
        static constexpr int size = 1, maxaxis = -1, nf = 0;
        template <int B, class P, bool A0>
        __device__ __forceinline__ static double eval(const ExprArgs&, int i, int j, int k) {
            return ((i + j + k) & 1) == C ? 1.0 : 0.0;
        }
        template <int B, class P, bool A0, int DI, int DJ, int DK, class Cx>
        __device__ __forceinline__ static double ev(const Cx& c) {
            return ((c.i0 + DI + c.j + DJ + c.k + DK) & 1) == C ? 1.0 : 0.0;
        }
        template <bool A0, class TS>
        static constexpr void taps(TS&, const TapGrid&) {}
    };

    // ------------------------------------------------------------------------------------------- point-wise
    // BinOpDefMacros.hpp.in:15-17 (operand order preserved), AMDS.hpp:34-91, MinMax.hpp:51-52, Compare.hpp, Boolean.hpp
#define OPF_BINOP(Name, EXPR)                                                                                          \
    template <class L, class R>                                                                                        \
    struct Name {                                                                                                      \
        static constexpr int size = 1 + L::size + R::size;                                                             \
        static constexpr int maxaxis = L::maxaxis > R::maxaxis ? L::maxaxis : R::maxaxis;                              \
        static constexpr int nf = L::nf > R::nf ? L::nf : R::nf;                                                       \
        template <class P>                                                                                             \
        __device__ __forceinline__ static double math(double x, double y) {                                            \
            return EXPR;                                                                                               \
        }                                                                                                              \
        template <int B, class P, bool A0>                                                                             \
        __device__ __forceinline__ static double eval(const ExprArgs& a, int i, int j, int k) {                        \
            const double x = L::template eval<B + 1, P, A0>(a, i, j, k);                                               \
            const double y = R::template eval<B + 1 + L::size, P, A0>(a, i, j, k);                                     \
            return math<P>(x, y);                                                                                      \
        }                                                                                                              \
        template <int B, class P, bool A0, int DI, int DJ, int DK, class C>                                            \
        __device__ __forceinline__ static double ev(const C& c) {                                                      \
            const double x = L::template ev<B + 1, P, A0, DI, DJ, DK>(c);                                              \
            const double y = R::template ev<B + 1 + L::size, P, A0, DI, DJ, DK>(c);                                    \
            return math<P>(x, y);                                                                                      \
        }                                                                                                              \
        template <bool A0, class TS>                                                                                   \
        static constexpr void taps(TS& ts, const TapGrid& in) {                                                        \
            L::template taps<A0>(ts, in);                                                                              \
            R::template taps<A0>(ts, in);                                                                              \
        }                                                                                                              \
    };
    OPF_BINOP(Add, P::add(x, y))
    OPF_BINOP(Sub, P::sub(x, y))
    OPF_BINOP(Mul, P::mul(x, y))
    OPF_BINOP(Div, P::div(x, y))
    OPF_BINOP(Min, (y < x ? y : x))// std::min(a,b) = (b<a)?b:a
    OPF_BINOP(Max, (x < y ? y : x))// std::max(a,b) = (a<b)?b:a
    OPF_BINOP(Pow, pow(x, y))
    OPF_BINOP(Lt, (x < y ? 1.0 : 0.0))
    OPF_BINOP(Le, (x <= y ? 1.0 : 0.0))
    OPF_BINOP(Gt, (x > y ? 1.0 : 0.0))
    OPF_BINOP(Ge, (x >= y ? 1.0 : 0.0))
    OPF_BINOP(Eq, (x == y ? 1.0 : 0.0))
    OPF_BINOP(Ne, (x != y ? 1.0 : 0.0))
    OPF_BINOP(And, ((x != 0.0 && y != 0.0) ? 1.0 : 0.0))
    OPF_BINOP(Or, ((x != 0.0 || y != 0.0) ? 1.0 : 0.0))
    // the remaining binary math functors of AMDS.hpp:40-51 (CUDA's double-precision libm; integer second operands are cast like the
    // implicit conversion of the reference's call)
    OPF_BINOP(FMod, fmod(x, y))
    OPF_BINOP(Remainder, remainder(x, y))
    OPF_BINOP(FDim, fdim(x, y))
    OPF_BINOP(Hypot, hypot(x, y))
    OPF_BINOP(ATan2, atan2(x, y))
    OPF_BINOP(Ldexp, ldexp(x, (int) y))
    OPF_BINOP(Scalbn, scalbn(x, (int) y))
    OPF_BINOP(Scalbln, scalbn(x, (int) y))
    OPF_BINOP(Nextafter, nextafter(x, y))
    OPF_BINOP(Nexttoward, nextafter(x, y))
    OPF_BINOP(Copysing, copysign(x, y))
#undef OPF_BINOP

#define OPF_UNIOP(Name, EXPR)                                                                                          \
    template <class E>                                                                                                 \
    struct Name {                                                                                                      \
        static constexpr int size = 1 + E::size, maxaxis = E::maxaxis, nf = E::nf;                                     \
        template <class P>                                                                                             \
        __device__ __forceinline__ static double math(double x) {                                                      \
            return EXPR;                                                                                               \
        }                                                                                                              \
        template <int B, class P, bool A0>                                                                             \
        __device__ __forceinline__ static double eval(const ExprArgs& a, int i, int j, int k) {                        \
            return math<P>(E::template eval<B + 1, P, A0>(a, i, j, k));                                                \
        }                                                                                                              \
        template <int B, class P, bool A0, int DI, int DJ, int DK, class C>                                            \
        __device__ __forceinline__ static double ev(const C& c) {                                                      \
            return math<P>(E::template ev<B + 1, P, A0, DI, DJ, DK>(c));                                               \
        }                                                                                                              \
        template <bool A0, class TS>                                                                                   \
        static constexpr void taps(TS& ts, const TapGrid& in) {                                                        \
            E::template taps<A0>(ts, in);                                                                              \
        }                                                                                                              \
    };
    OPF_UNIOP(Neg, -x)
    OPF_UNIOP(Pos, x)
    OPF_UNIOP(Not, (x == 0.0 ? 1.0 : 0.0))
    OPF_UNIOP(Sqrt, sqrt(x))
    OPF_UNIOP(Abs, fabs(x))
    OPF_UNIOP(Exp, exp(x))
    OPF_UNIOP(Log, log(x))
    OPF_UNIOP(Sin, sin(x))
    OPF_UNIOP(Cos, cos(x))
    OPF_UNIOP(Tan, tan(x))
    OPF_UNIOP(Tanh, tanh(x))
    OPF_UNIOP(Pow2, P::mul(x, x))
    // the remaining unary math functors of AMDS.hpp:56-89 (integer-valued results are stored as Real, as an assignment to a
    // CartesianField<Real> does in the reference)
    OPF_UNIOP(Exp2, exp2(x))
    OPF_UNIOP(Expm1, expm1(x))
    OPF_UNIOP(Log10, log10(x))
    OPF_UNIOP(Log2, log2(x))
    OPF_UNIOP(Log1p, log1p(x))
    OPF_UNIOP(Cbrt, cbrt(x))
    OPF_UNIOP(ASin, asin(x))
    OPF_UNIOP(ACos, acos(x))
    OPF_UNIOP(ATan, atan(x))
    OPF_UNIOP(Sinh, sinh(x))
    OPF_UNIOP(Cosh, cosh(x))
    OPF_UNIOP(ASinh, asinh(x))
    OPF_UNIOP(ACosh, acosh(x))
    OPF_UNIOP(ATanh, atanh(x))
    OPF_UNIOP(Erf, erf(x))
    OPF_UNIOP(Erfc, erfc(x))
    OPF_UNIOP(TGamma, tgamma(x))
    OPF_UNIOP(LGamma, lgamma(x))
    OPF_UNIOP(Ceil, ceil(x))
    OPF_UNIOP(Floor, floor(x))
    OPF_UNIOP(Trunc, trunc(x))
    OPF_UNIOP(Round, round(x))
    OPF_UNIOP(LRound, (double) lround(x))
    OPF_UNIOP(LLRound, (double) llround(x))
    OPF_UNIOP(NearbyInt, nearbyint(x))
    OPF_UNIOP(Rint, rint(x))
    OPF_UNIOP(LRint, (double) lrint(x))
    OPF_UNIOP(LLRint, (double) llrint(x))
    OPF_UNIOP(ILogb, (double) ilogb(x))
    OPF_UNIOP(Logb, logb(x))
#undef OPF_UNIOP

    // UniOpAdaptor<functor> / BinOpAdaptor<functor> (src/Core/Operator/PerElemOpAdaptor.hpp:23-110): a user functor applied element by
    // element.  Fn is the functor's TYPE (a NamedFunctor is a stateless constexpr object): it must be default-constructible and its
    // call operator usable in device code -- `__host__ __device__`, or constexpr under nvcc's --expt-relaxed-constexpr.
    template <class Fn, class E>
    struct Adapt1 {
        static constexpr int size = 1 + E::size, maxaxis = E::maxaxis, nf = E::nf;
        template <int B, class P, bool A0>
        __device__ __forceinline__ static double eval(const ExprArgs& a, int i, int j, int k) {
            return static_cast<double>(Fn {}(E::template eval<B + 1, P, A0>(a, i, j, k)));
        }
        template <int B, class P, bool A0, int DI, int DJ, int DK, class C>
        __device__ __forceinline__ static double ev(const C& c) {
            return static_cast<double>(Fn {}(E::template ev<B + 1, P, A0, DI, DJ, DK>(c)));
        }
        template <bool A0, class TS>
        static constexpr void taps(TS& ts, const TapGrid& in) {
            E::template taps<A0>(ts, in);
        }
    };
    template <class Fn, class L, class R>
    struct Adapt2 {
        static constexpr int size = 1 + L::size + R::size;
        static constexpr int maxaxis = L::maxaxis > R::maxaxis ? L::maxaxis : R::maxaxis;
        static constexpr int nf = L::nf > R::nf ? L::nf : R::nf;
        template <int B, class P, bool A0>
        __device__ __forceinline__ static double eval(const ExprArgs& a, int i, int j, int k) {
            return static_cast<double>(Fn {}(L::template eval<B + 1, P, A0>(a, i, j, k), R::template eval<B + 1 + L::size, P, A0>(a, i, j, k)));
        }
        template <int B, class P, bool A0, int DI, int DJ, int DK, class C>
        __device__ __forceinline__ static double ev(const C& c) {
            return static_cast<double>(Fn {}(L::template ev<B + 1, P, A0, DI, DJ, DK>(c), R::template ev<B + 1 + L::size, P, A0, DI, DJ, DK>(c)));
        }
        template <bool A0, class TS>
        static constexpr void taps(TS& ts, const TapGrid& in) {
            L::template taps<A0>(ts, in);
            R::template taps<A0>(ts, in);
        }
    };

    // CondOp::eval (Conditional.hpp:37-40)
    template <class Cc, class A, class Bb>
    struct Cond {
        static constexpr int size = 1 + Cc::size + A::size + Bb::size;
        static constexpr int maxaxis = (Cc::maxaxis > A::maxaxis ? Cc::maxaxis : A::maxaxis) > Bb::maxaxis
                                               ? (Cc::maxaxis > A::maxaxis ? Cc::maxaxis : A::maxaxis)
                                               : Bb::maxaxis;
        static constexpr int nf = (Cc::nf > A::nf ? Cc::nf : A::nf) > Bb::nf ? (Cc::nf > A::nf ? Cc::nf : A::nf) : Bb::nf;
        template <int B, class P, bool A0>
        __device__ __forceinline__ static double eval(const ExprArgs& a, int i, int j, int k) {
            const double c = Cc::template eval<B + 1, P, A0>(a, i, j, k);
            return c != 0.0 ? A::template eval<B + 1 + Cc::size, P, A0>(a, i, j, k)
                            : Bb::template eval<B + 1 + Cc::size + A::size, P, A0>(a, i, j, k);
        }
        template <int B, class P, bool A0, int DI, int DJ, int DK, class C>
        __device__ __forceinline__ static double ev(const C& c) {
            const double cv = Cc::template ev<B + 1, P, A0, DI, DJ, DK>(c);
            return cv != 0.0 ? A::template ev<B + 1 + Cc::size, P, A0, DI, DJ, DK>(c)
                             : Bb::template ev<B + 1 + Cc::size + A::size, P, A0, DI, DJ, DK>(c);
        }
        template <bool A0, class TS>
        static constexpr void taps(TS& ts, const TapGrid& in) {
            Cc::template taps<A0>(ts, in);
            A::template taps<A0>(ts, in);
            Bb::template taps<A0>(ts, in);
        }
    };

    // ------------------------------------------------------------------------------------------- stencils
    // helper: boilerplate shared by all 1-operand stencil nodes.  TAP(n) evaluates the operand at offset n along axis D.
#define OPF_STENCIL_HEAD(LO, HI)                                                                                       \
    static constexpr int size = 1 + E::size, maxaxis = (D > E::maxaxis ? D : E::maxaxis), nf = E::nf;                  \
    template <bool A0, class TS>                                                                                       \
    static constexpr void taps(TS& ts, const TapGrid& in) {                                                            \
        E::template taps<A0>(ts, tap_shift(in, D, LO, HI));                                                            \
    }
#define OPF_EVAL_BEGIN                                                                                                 \
    template <int B, class P, bool A0>                                                                                 \
    __device__ __forceinline__ static double eval(const ExprArgs& a, int i, int j, int k) {                            \
        const GAcc acc{a.ax[D], axis_of<D>(i, j, k)};                                                                  \
        auto TAP = [&](auto n) { return E::template eval<B + 1, P, A0>(a, OPF_SHIFT(D, decltype(n)::value)); };
#define OPF_EV_BEGIN                                                                                                   \
    template <int B, class P, bool A0, int DI, int DJ, int DK, class C>                                                \
    __device__ __forceinline__ static double ev(const C& c) {                                                          \
        const ExprArgs& a = c.a;                                                                                       \
        const WAcc<C, D, (D == 0 ? DI : (D == 1 ? DJ : DK))> acc{c};                                                   \
        auto TAP = [&](auto n) {                                                                                       \
            constexpr int o = decltype(n)::value;                                                                      \
            return E::template ev<B + 1, P, A0, DI + (D == 0 ? o : 0), DJ + (D == 1 ? o : 0), DK + (D == 2 ? o : 0)>(c); \
        };
    // mesh-coefficient accessors: O is the compile-time offset from the stencil's own index q along its axis
    enum { CF_X = 0, CF_DX = 1, CF_RDX = 2, CF_RDXH = 3, CF_RDXC = 4, CF_RDX2 = 5 };
    struct GAcc {// direct global loads (run-time coordinate evaluator)
        static constexpr bool uni = false;
        const AxisView& ax;
        int q;
        // (serving a single-spacing axis from the kernel-parameter constant bank behind a warp-uniform `ax.uniform ? ... : ...` was tried
        // here: the two-way code made fe_tg3d 35 % larger and its momentum operators 15 % slower -- profiles/r2_summary.md section 6)
        template <int O> __device__ __forceinline__ double x() const { return __ldg(ax.x + q + O); }
        template <int O> __device__ __forceinline__ double dx() const { return __ldg(ax.dx + q + O); }
        template <int O> __device__ __forceinline__ double rdx() const { return __ldg(ax.rdx + q + O); }
        template <int O> __device__ __forceinline__ double rdxh() const { return __ldg(ax.rdxh + q + O); }
        template <int O> __device__ __forceinline__ double rdxc() const { return __ldg(ax.rdxc + q + O); }
    };
    template <class C, int D, int QO>
    struct WAcc {// register-cached coefficients of the window context (hoisted out of the march loop)
        static constexpr bool uni = C::UNI;
        const C& c;
        __device__ __forceinline__ double rdx2() const { return c.a.ax[D].u[CF_RDX2]; }
        template <int O> __device__ __forceinline__ double x() const { return c.template coef<D, CF_X, QO + O>(); }
        template <int O> __device__ __forceinline__ double dx() const { return c.template coef<D, CF_DX, QO + O>(); }
        template <int O> __device__ __forceinline__ double rdx() const { return c.template coef<D, CF_RDX, QO + O>(); }
        template <int O> __device__ __forceinline__ double rdxh() const { return c.template coef<D, CF_RDXH, QO + O>(); }
        template <int O> __device__ __forceinline__ double rdxc() const { return c.template coef<D, CF_RDXC, QO + O>(); }
    };

    // D2SecondOrderCentered<d>::eval (D2SecondOrderCentered.hpp:161-171)
    template <class P, class Acc>
    __device__ __forceinline__ double d2c_math(double l, double c, double r, bool center, const Acc& m) {
        if constexpr (P::fast && Acc::uni) {
            // uniform axis: dx_l == dx_r == dx_c == dx for Corner and Center alike -> one multiply by 1/dx^2
            return ((r - c) - (c - l)) * m.rdx2();
        } else if constexpr (P::fast) {
            if (!center) return ((r - c) * m.template rdx<0>() - (c - l) * m.template rdx<-1>()) * m.template rdxh<0>();
            return ((r - c) * m.template rdxh<1>() - (c - l) * m.template rdxh<0>()) * m.template rdxc<0>();
        } else {
            double dxl, dxr;
            if (!center) {
                dxl = m.template dx<-1>();
                dxr = m.template dx<0>();
            } else {
                dxl = P::mul(P::add(m.template dx<-1>(), m.template dx<0>()), 0.5);
                dxr = P::mul(P::add(m.template dx<0>(), m.template dx<1>()), 0.5);
            }
            const double dxc = P::mul(P::add(dxl, dxr), 0.5);
            return P::div(P::sub(P::div(P::sub(r, c), dxr), P::div(P::sub(c, l), dxl)), dxc);
        }
    }
    template <int D, class E>
    struct D2C {
        OPF_STENCIL_HEAD(-1, 1)
        OPF_EVAL_BEGIN
            return d2c_math<P>(TAP(IC<-1>{}), TAP(IC<0>{}), TAP(IC<1>{}), (a.loc[B] >> D) & 1, acc);
        }
        OPF_EV_BEGIN
            return d2c_math<P>(TAP(IC<-1>{}), TAP(IC<0>{}), TAP(IC<1>{}), (a.loc[B] >> D) & 1, acc);
        }
    };

    // D1FirstOrderCentered<d>::eval (D1FirstOrderCentered.hpp:31-36); result loc flipped in prepare (:46)
    // Fast arithmetic: the divides become multiplies by the per-axis reciprocal arrays (rdxh[i] = 1 / ((dx[i-1] + dx[i]) / 2))
    template <class P, class Acc>
    __device__ __forceinline__ double d1c_center(double e0, double em1, const Acc& m) {// Center -> Corner
        if constexpr (P::fast) return (e0 - em1) * m.template rdxh<0>();
        else return P::mul(P::div(P::sub(e0, em1), P::add(m.template dx<-1>(), m.template dx<0>())), 2.0);
    }
    template <class P, class Acc>
    __device__ __forceinline__ double d1c_corner(double ep1, double e0, const Acc& m) {// Corner -> Center
        if constexpr (P::fast) return (ep1 - e0) * m.template rdx<0>();
        else return P::div(P::sub(ep1, e0), m.template dx<0>());
    }
    template <int D, class E>
    struct D1C {
        OPF_STENCIL_HEAD(-1, 1)
        OPF_EVAL_BEGIN
            if ((a.loc[B] >> D) & 1) return d1c_center<P>(TAP(IC<0>{}), TAP(IC<-1>{}), acc);
            return d1c_corner<P>(TAP(IC<1>{}), TAP(IC<0>{}), acc);
        }
        OPF_EV_BEGIN
            if ((a.loc[B] >> D) & 1) return d1c_center<P>(TAP(IC<0>{}), TAP(IC<-1>{}), acc);
            return d1c_corner<P>(TAP(IC<1>{}), TAP(IC<0>{}), acc);
        }
    };

    // D1FirstOrderBiasedDownwind<d>::eval (D1FirstOrderBiasedDownwind.hpp:53-57)
    template <class P, class Acc>
    __device__ __forceinline__ double d1dn_math(double e0, double em1, bool center, const Acc& m) {
        if constexpr (P::fast) return (e0 - em1) * (center ? m.template rdxh<0>() : m.template rdx<-1>());
        else {
            const double h = center ? P::mul(P::add(m.template dx<-1>(), m.template dx<0>()), 0.5) : m.template dx<-1>();
            return P::div(P::sub(e0, em1), h);
        }
    }
    template <class P, class Acc>
    __device__ __forceinline__ double d1up_math(double ep1, double e0, bool center, const Acc& m) {
        if constexpr (P::fast) return (ep1 - e0) * (center ? m.template rdxh<1>() : m.template rdx<0>());
        else {
            const double h = center ? P::mul(P::add(m.template dx<0>(), m.template dx<1>()), 0.5) : m.template dx<0>();
            return P::div(P::sub(ep1, e0), h);
        }
    }
    template <int D, class E>
    struct D1Dn {
        OPF_STENCIL_HEAD(-1, 0)
        OPF_EVAL_BEGIN
            return d1dn_math<P>(TAP(IC<0>{}), TAP(IC<-1>{}), (a.loc[B] >> D) & 1, acc);
        }
        OPF_EV_BEGIN
            return d1dn_math<P>(TAP(IC<0>{}), TAP(IC<-1>{}), (a.loc[B] >> D) & 1, acc);
        }
    };
    // D1FirstOrderBiasedUpwind<d>::eval (D1FirstOrderBiasedUpwind.hpp:54-58)
    template <int D, class E>
    struct D1Up {
        OPF_STENCIL_HEAD(0, 1)
        OPF_EVAL_BEGIN
            return d1up_math<P>(TAP(IC<1>{}), TAP(IC<0>{}), (a.loc[B] >> D) & 1, acc);
        }
        OPF_EV_BEGIN
            return d1up_math<P>(TAP(IC<1>{}), TAP(IC<0>{}), (a.loc[B] >> D) & 1, acc);
        }
    };

    // D1WENO53{Down,Up}wind<d>::kernel (D1WENO53Downwind.hpp:136-151, D1WENO53Upwind.hpp:133-151): identical body,
    // fed with differently ordered one-sided differences d1..d5.
    template <class P>
    __device__ __forceinline__ double weno53_core(double d1, double d2, double d3, double d4, double d5) {
        if constexpr (P::fast) {
            const double c3 = 1. / 3., c6 = 1. / 6.;
            const double ddx1 = d1 * c3 - 7. * d2 * c6 + 11. * d3 * c6;
            const double ddx2 = -d2 * c6 + 5. * d3 * c6 + d4 * c3;
            const double ddx3 = d3 * c3 + 5. * d4 * c6 - d5 * c6;
            const double t1 = d1 - 2 * d2 + d3, t2 = d1 - 4 * d2 + 3 * d3;
            const double t3 = d2 - 2 * d3 + d4, t4 = d2 - d4;
            const double t5 = d3 - 2 * d4 + d5, t6 = 3 * d3 - 4 * d4 + d5;
            const double s1 = 13. / 12. * t1 * t1 + t2 * t2 * 0.25;
            const double s2 = 13. / 12. * t3 * t3 + t4 * t4 * 0.25;
            const double s3 = 13. / 12. * t5 * t5 + t6 * t6 * 0.25;
            const double eps = 1e-6 * fmax(fmax(fmax(d1 * d1, d2 * d2), fmax(d3 * d3, d4 * d4)), d5 * d5) + 1e-99;
            const double e1 = s1 + eps, e2 = s2 + eps, e3 = s3 + eps;
            // w_k = (g_k / e_k^2) / sum_j (g_j / e_j^2), with every a_k scaled by the common factor (e_2 e_3 / e_1)^2 * e_1^2 so that only TWO
            // FP64 reciprocals remain (each costs ~12 FP64-pipe instructions; the kernel is FP64-pipe bound, ncu: 71 % pipe active):
            //   f_k = e_k / e_1 ;  a_1 = .1 f_2^2 f_3^2 ;  a_2 = .6 f_3^2 ;  a_3 = .3 f_2^2
            // e_k can be ~1e-99 on flat data, so raw cross products e_j^2 e_k^2 would underflow; the RATIOS cannot misbehave:
            // eps <= e_k <= 33 max(d^2) + eps and eps >= 1e-6 max(d^2) bound e_k / e_1 by 3.3e7, f^2 f^2 by 1.2e30.
            const double r1 = 1. / e1;
            const double f2 = e2 * r1, f3 = e3 * r1;
            const double g2 = f2 * f2, g3 = f3 * f3;
            const double a1 = .1 * g2 * g3, a2 = .6 * g3, a3 = .3 * g2;
            const double inv = 1. / (a1 + a2 + a3);
            return (a1 * ddx1 + a2 * ddx2 + a3 * ddx3) * inv;
        } else {
            const double ddx1 = P::add(P::sub(P::div(d1, 3.), P::div(P::mul(7., d2), 6.)), P::div(P::mul(11., d3), 6.));
            const double ddx2 = P::add(P::add(P::div(-d2, 6.), P::div(P::mul(5., d3), 6.)), P::div(d4, 3.));
            const double ddx3 = P::sub(P::add(P::div(d3, 3.), P::div(P::mul(5., d4), 6.)), P::div(d5, 6.));
            const double t1 = P::add(P::sub(d1, P::mul(2., d2)), d3);
            const double t2 = P::add(P::sub(d1, P::mul(4., d2)), P::mul(3., d3));
            const double t3 = P::add(P::sub(d2, P::mul(2., d3)), d4);
            const double t4 = P::sub(d2, d4);
            const double t5 = P::add(P::sub(d3, P::mul(2., d4)), d5);
            const double t6 = P::add(P::sub(P::mul(3., d3), P::mul(4., d4)), d5);
            const double k1312 = 13. / 12.;
            const double s1 = P::add(P::mul(k1312, P::mul(t1, t1)), P::div(P::mul(t2, t2), 4.));
            const double s2 = P::add(P::mul(k1312, P::mul(t3, t3)), P::div(P::mul(t4, t4), 4.));
            const double s3 = P::add(P::mul(k1312, P::mul(t5, t5)), P::div(P::mul(t6, t6), 4.));
            // std::max over an initializer_list == left fold of std::max
            double m = P::mul(d1, d1);
            const double q2 = P::mul(d2, d2), q3 = P::mul(d3, d3), q4 = P::mul(d4, d4), q5 = P::mul(d5, d5);
            m = m < q2 ? q2 : m;
            m = m < q3 ? q3 : m;
            m = m < q4 ? q4 : m;
            m = m < q5 ? q5 : m;
            const double eps = P::add(P::mul(1e-6, m), 1e-99);
            const double e1 = P::add(s1, eps), e2 = P::add(s2, eps), e3 = P::add(s3, eps);
            const double a1 = P::div(.1, P::mul(e1, e1));
            const double a2 = P::div(.6, P::mul(e2, e2));
            const double a3 = P::div(.3, P::mul(e3, e3));
            const double sum = P::add(P::add(a1, a2), a3);
            const double w1 = P::div(a1, sum), w2 = P::div(a2, sum), w3 = P::div(a3, sum);
            return P::add(P::add(P::mul(w1, ddx1), P::mul(w2, ddx2)), P::mul(w3, ddx3));
        }
    }
    // one-sided differences divided by h = dx(d,i) ("uniform mesh is assumed", D1WENO53Upwind.hpp:76)
    template <class P, class Acc>
    __device__ __forceinline__ double weno53(double a0, double a1, double a2, double a3, double a4, double a5, const Acc& m) {
        if constexpr (P::fast) {
            const double rh = m.template rdx<0>();
            return weno53_core<P>((a1 - a0) * rh, (a2 - a1) * rh, (a3 - a2) * rh, (a4 - a3) * rh, (a5 - a4) * rh);
        } else {
            const double h = m.template dx<0>();
            return weno53_core<P>(P::div(P::sub(a1, a0), h), P::div(P::sub(a2, a1), h), P::div(P::sub(a3, a2), h),
                                  P::div(P::sub(a4, a3), h), P::div(P::sub(a5, a4), h));
        }
    }
    // D1WENO53Downwind<d>::eval (D1WENO53Downwind.hpp:77-86): taps i-3..i+2;  d1=(pm2-pm3)/h ... d5=(pp2-pp1)/h
    template <int D, class E>
    struct WenoDn {
        OPF_STENCIL_HEAD(-3, 2)
        OPF_EVAL_BEGIN
            return weno53<P>(TAP(IC<-3>{}), TAP(IC<-2>{}), TAP(IC<-1>{}), TAP(IC<0>{}), TAP(IC<1>{}), TAP(IC<2>{}), acc);
        }
        OPF_EV_BEGIN
            return weno53<P>(TAP(IC<-3>{}), TAP(IC<-2>{}), TAP(IC<-1>{}), TAP(IC<0>{}), TAP(IC<1>{}), TAP(IC<2>{}), acc);
        }
    };
    // D1WENO53Upwind<d>::eval (D1WENO53Upwind.hpp:75-84): taps i-2..i+3;  d1=(pp3-pp2)/h ... d5=(pm1-pm2)/h, i.e. the
    // same differences with the sequence reversed and negated twice: d_k = (b_k - b_{k+1})/h with b = (pp3,...,pm2)
    template <class P, class Acc>
    __device__ __forceinline__ double weno53_up(double pm2, double pm1, double p0, double pp1, double pp2, double pp3, const Acc& m) {
        if constexpr (P::fast) {
            const double rh = m.template rdx<0>();
            return weno53_core<P>((pp3 - pp2) * rh, (pp2 - pp1) * rh, (pp1 - p0) * rh, (p0 - pm1) * rh, (pm1 - pm2) * rh);
        } else {
            const double h = m.template dx<0>();
            return weno53_core<P>(P::div(P::sub(pp3, pp2), h), P::div(P::sub(pp2, pp1), h), P::div(P::sub(pp1, p0), h),
                                  P::div(P::sub(p0, pm1), h), P::div(P::sub(pm1, pm2), h));
        }
    }
    template <int D, class E>
    struct WenoUp {
        OPF_STENCIL_HEAD(-2, 3)
        OPF_EVAL_BEGIN
            return weno53_up<P>(TAP(IC<-2>{}), TAP(IC<-1>{}), TAP(IC<0>{}), TAP(IC<1>{}), TAP(IC<2>{}), TAP(IC<3>{}), acc);
        }
        OPF_EV_BEGIN
            return weno53_up<P>(TAP(IC<-2>{}), TAP(IC<-1>{}), TAP(IC<0>{}), TAP(IC<1>{}), TAP(IC<2>{}), TAP(IC<3>{}), acc);
        }
    };

    // D1Linear<d, Cen2Cor>::eval (D1Linear.hpp:36-42) with Interpolator1D::intp (Interpolator.hpp:21-23,38-40)
    template <class P, class Acc>
    __device__ __forceinline__ double intp_c2n(double y1, double y2, const Acc& m) {
        if constexpr (P::fast) {
            // x1 - x = -dx[-1] / 2 and x2 - x = +dx[0] / 2, so the interpolant is (dx[-1] y2 + dx[0] y1) / (dx[-1] + dx[0]): no divide,
            // no coordinate loads (on a uniform axis: (y1 + y2) / 2 up to one rounding)
            return (m.template dx<-1>() * y2 + m.template dx<0>() * y1) * (0.5 * m.template rdxh<0>());
        }
        const double x1 = P::add(m.template x<-1>(), P::mul(0.5, m.template dx<-1>()));
        const double x2 = P::add(m.template x<0>(), P::mul(0.5, m.template dx<0>()));
        const double x = m.template x<0>();
        const double u1 = P::sub(x1, x), u2 = P::sub(x2, x);
        return P::div(P::sub(P::mul(u1, y2), P::mul(u2, y1)), P::sub(u1, u2));
    }
    template <int D, class E>
    struct IntpC2N {
        OPF_STENCIL_HEAD(-1, 0)
        OPF_EVAL_BEGIN
            return intp_c2n<P>(TAP(IC<-1>{}), TAP(IC<0>{}), acc);
        }
        OPF_EV_BEGIN
            return intp_c2n<P>(TAP(IC<-1>{}), TAP(IC<0>{}), acc);
        }
    };
    // D1Linear<d, Cor2Cen>::eval (D1Linear.hpp:44): Math::mid (Interpolator.hpp:83)
    template <int D, class E>
    struct IntpN2C {
        OPF_STENCIL_HEAD(0, 1)
        OPF_EVAL_BEGIN
            (void) acc;
            return P::mul(P::add(TAP(IC<0>{}), TAP(IC<1>{})), 0.5);
        }
        OPF_EV_BEGIN
            (void) acc;
            (void) a;
            return P::mul(P::add(TAP(IC<0>{}), TAP(IC<1>{})), 0.5);
        }
    };
#undef OPF_STENCIL_HEAD
#undef OPF_EVAL_BEGIN
#undef OPF_EV_BEGIN

    // ------------------------------------------------------------------------------------------- flux-limiter interpolators
    // D1FluxLimiterImpl<Kernel, d, dir>::eval(u, e, i) (src/Core/Operator/Interpolator/D1FluxLimiter.hpp:41-190): the face value of e
    // reconstructed from the upwind side, the side chosen per cell by the sign of the advecting field u:
    //   u[i] > 0 ? Upwind(e, i) : Downwind(e, i)
    // Cen2Cor (DIR 0) taps e at i-2 .. i+1 and yields a Corner value; Cor2Cen (DIR 1) taps i-1 .. i+2 and yields a Center value.
    // "Linear" kernels (the kappa family, FluxLimiterKernels.hpp:32-52) combine the two slopes directly; the others limit through
    // r = slope_f / (slope_u + 1e-16).  Operation order follows the reference so that Exact arithmetic is bit-identical.
    template <int NUM, int DEN>
    struct FlKappa {// KappaKernel<s>::eval(slop_u, slop_f): (1 + kappa) / 2. * slop_f + (1 - kappa) / 2. * slop_u
        static constexpr bool linear = true;
        template <class P>
        __device__ __forceinline__ static double eval2(double su, double sf) {
            constexpr double kappa = (double) NUM / (double) DEN;
            return P::add(P::mul((1 + kappa) / 2., sf), P::mul((1 - kappa) / 2., su));
        }
    };
    __device__ __forceinline__ double fl_max(double a, double b) { return a < b ? b : a; }// std::max
    __device__ __forceinline__ double fl_min(double a, double b) { return b < a ? b : a; }// std::min
    struct FlMinmodK {// std::max(0., std::min(r, 1.))
        static constexpr bool linear = false;
        template <class P>
        __device__ __forceinline__ static double eval1(double r) { return fl_max(0., fl_min(r, 1.)); }
    };
    struct FlSuperbeeK {// std::max({0., std::min(2. * r, 1.), std::min(r, 2.)})
        static constexpr bool linear = false;
        template <class P>
        __device__ __forceinline__ static double eval1(double r) { return fl_max(fl_max(0., fl_min(P::mul(2., r), 1.)), fl_min(r, 2.)); }
    };
    struct FlMusclK {// std::max(0., std::min({2 * r, (r + 1) / 2., 2.}))
        static constexpr bool linear = false;
        template <class P>
        __device__ __forceinline__ static double eval1(double r) { return fl_max(0., fl_min(fl_min(P::mul(2., r), P::div(P::add(r, 1.), 2.)), 2.)); }
    };
    struct FlHarmonicK {// (r + std::fabs(r)) / (r + 1)
        static constexpr bool linear = false;
        template <class P>
        __device__ __forceinline__ static double eval1(double r) { return P::div(P::add(r, fabs(r)), P::add(r, 1.)); }
    };
    struct FlAlbadaK {// r * (r + 1) / (r * r + 1)
        static constexpr bool linear = false;
        template <class P>
        __device__ __forceinline__ static double eval1(double r) { return P::div(P::mul(r, P::add(r, 1.)), P::add(P::mul(r, r), 1.)); }
    };
    // one side: y2 +- dx * 0.5 * K(...) with slopes (y2 - y1) / (x2 - x1) and (y3 - y2) / (x3 - x2) assigned to upwind / far
    template <class K, class P>
    __device__ __forceinline__ double fl_side(double su, double sf, double y2, double h, bool plus) {
        double t;
        if constexpr (K::linear) t = P::mul(P::mul(h, 0.5), K::template eval2<P>(su, sf));
        else {
            const double r = P::div(sf, P::add(su, 1e-16));
            t = P::mul(P::mul(P::mul(h, 0.5), K::template eval1<P>(r)), su);
        }
        return plus ? P::add(y2, t) : P::sub(y2, t);
    }
    // q0..q3: operand at (C2N) i-2, i-1, i, i+1 / (N2C) i-1, i, i+1, i+2
    template <class K, int DIR, class P, class Acc>
    __device__ __forceinline__ double fl_math(double uv, double q0, double q1, double q2, double q3, const Acc& m) {
        if constexpr (DIR == 0) {
            if (uv > 0.) {// D1FluxLimiterUpwindImpl<.., Cen2Cor> :46-63
                const double x1 = P::add(m.template x<-2>(), P::mul(0.5, m.template dx<-2>()));
                const double x2 = P::add(m.template x<-1>(), P::mul(0.5, m.template dx<-1>()));
                const double x3 = P::add(m.template x<0>(), P::mul(0.5, m.template dx<0>()));
                const double su = P::div(P::sub(q1, q0), P::sub(x2, x1)), sf = P::div(P::sub(q2, q1), P::sub(x3, x2));
                return fl_side<K, P>(su, sf, q1, m.template dx<-1>(), true);
            }
            // D1FluxLimiterDownwindImpl<.., Cen2Cor> :95-112
            const double x1 = P::add(m.template x<-1>(), P::mul(m.template dx<-1>(), 0.5));
            const double x2 = P::add(m.template x<0>(), P::mul(m.template dx<0>(), 0.5));
            const double x3 = P::add(m.template x<1>(), P::mul(m.template dx<1>(), 0.5));
            const double su = P::div(P::sub(q3, q2), P::sub(x3, x2)), sf = P::div(P::sub(q2, q1), P::sub(x2, x1));
            return fl_side<K, P>(su, sf, q2, m.template dx<0>(), false);
        } else {
            if (uv > 0.) {// D1FluxLimiterUpwindImpl<.., Cor2Cen> :67-84
                const double x1 = m.template x<-1>(), x2 = m.template x<0>(), x3 = m.template x<1>();
                const double su = P::div(P::sub(q1, q0), P::sub(x2, x1)), sf = P::div(P::sub(q2, q1), P::sub(x3, x2));
                return fl_side<K, P>(su, sf, q1, m.template dx<0>(), true);
            }
            // D1FluxLimiterDownwindImpl<.., Cor2Cen> :116-133
            const double x1 = m.template x<0>(), x2 = m.template x<1>(), x3 = m.template x<2>();
            const double su = P::div(P::sub(q3, q2), P::sub(x3, x2)), sf = P::div(P::sub(q2, q1), P::sub(x2, x1));
            return fl_side<K, P>(su, sf, q2, m.template dx<0>(), false);
        }
    }
    template <class K, int DIR, int D, class U, class E>
    struct FluxLim {
        static constexpr int size = 1 + U::size + E::size;
        static constexpr int maxaxis = (D > U::maxaxis ? D : U::maxaxis) > E::maxaxis ? (D > U::maxaxis ? D : U::maxaxis) : E::maxaxis;
        static constexpr int nf = U::nf > E::nf ? U::nf : E::nf;
        static constexpr int LO = DIR == 0 ? -2 : -1;
        template <bool A0, class TS>
        static constexpr void taps(TS& ts, const TapGrid& in) {
            U::template taps<A0>(ts, in);
            E::template taps<A0>(ts, tap_shift(in, D, LO, LO + 3));
        }
        template <int B, class P, bool A0>
        __device__ __forceinline__ static double eval(const ExprArgs& a, int i, int j, int k) {
            const GAcc acc{a.ax[D], axis_of<D>(i, j, k)};
            constexpr int BE = B + 1 + U::size;
            const double uv = U::template eval<B + 1, P, A0>(a, i, j, k);
            return fl_math<K, DIR, P>(uv, E::template eval<BE, P, A0>(a, OPF_SHIFT(D, LO)), E::template eval<BE, P, A0>(a, OPF_SHIFT(D, LO + 1)),
                                      E::template eval<BE, P, A0>(a, OPF_SHIFT(D, LO + 2)), E::template eval<BE, P, A0>(a, OPF_SHIFT(D, LO + 3)), acc);
        }
        template <int B, class P, bool A0, int DI, int DJ, int DK, class C>
        __device__ __forceinline__ static double ev(const C& c) {
            const WAcc<C, D, (D == 0 ? DI : (D == 1 ? DJ : DK))> acc{c};
            constexpr int BE = B + 1 + U::size;
            const double uv = U::template ev<B + 1, P, A0, DI, DJ, DK>(c);
#define OPF_FL_TAP(o) E::template ev<BE, P, A0, DI + (D == 0 ? (o) : 0), DJ + (D == 1 ? (o) : 0), DK + (D == 2 ? (o) : 0)>(c)
            return fl_math<K, DIR, P>(uv, OPF_FL_TAP(LO), OPF_FL_TAP(LO + 1), OPF_FL_TAP(LO + 2), OPF_FL_TAP(LO + 3), acc);
#undef OPF_FL_TAP
        }
    };
    // named nodes of the signature grammar: Fl<Scheme><C2N|N2C><axis, u, e>   (D1FluxLimiterBasedIntpOp.hpp:22-61)
#define OPF_FLNODE(Name, ...)                                                                                          \
    template <int D, class U, class E>                                                                                 \
    struct Name##C2N : FluxLim<__VA_ARGS__, 0, D, U, E> {};                                                            \
    template <int D, class U, class E>                                                                                 \
    struct Name##N2C : FluxLim<__VA_ARGS__, 1, D, U, E> {};
    OPF_FLNODE(FlCentral, FlKappa<1, 1>)
    OPF_FLNODE(FlQuick, FlKappa<1, 2>)
    OPF_FLNODE(FlCui, FlKappa<1, 3>)
    OPF_FLNODE(FlFromm, FlKappa<0, 1>)
    OPF_FLNODE(FlLui, FlKappa<-1, 1>)
    OPF_FLNODE(FlMinmod, FlMinmodK)
    OPF_FLNODE(FlSuperbee, FlSuperbeeK)
    OPF_FLNODE(FlMuscl, FlMusclK)
    OPF_FLNODE(FlHarmonic, FlHarmonicK)
    OPF_FLNODE(FlAlbada, FlAlbadaK)
#undef OPF_FLNODE

    // Convolution<n0, n1[, n2]>::eval (src/Core/Operator/Convolution/Convolution.hpp:45-63): sum over the kernel box of
    // kernel[idx - i + n/2] * e[idx], accumulated by rangeReduce_s in x-fastest order starting from 0.  The kernel tensor is a
    // ScalarExpr operand: its entries travel in consecutive scalar slots S<K0>, S<K0+1>, ... (x-fastest like FixedSizeTensor).
    template <int N0, int N1, int N2, int K0, class E>
    struct Conv {
        static constexpr int size = 1 + E::size, nf = E::nf;
        static constexpr int maxaxis = (N2 > 1 ? 2 : (N1 > 1 ? 1 : 0)) > E::maxaxis ? (N2 > 1 ? 2 : (N1 > 1 ? 1 : 0)) : E::maxaxis;
        static constexpr int H0 = N0 / 2, H1 = N1 / 2, H2 = N2 / 2;
        template <bool A0, class TS>
        static constexpr void taps(TS& ts, const TapGrid& in) {
            E::template taps<A0>(ts, tap_shift(tap_shift(tap_shift(in, 0, -H0, H0), 1, -H1, H1), 2, -H2, H2));
        }
        template <int B, class P, bool A0>
        __device__ __forceinline__ static double eval(const ExprArgs& a, int i, int j, int k) {
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < N2; ++c)
#pragma unroll
                for (int b = 0; b < N1; ++b)
#pragma unroll
                    for (int x = 0; x < N0; ++x)
                        acc = P::add(acc, P::mul(a.s[K0 + x + N0 * (b + N1 * c)], E::template eval<B + 1, P, A0>(a, i + x - H0, j + b - H1, k + c - H2)));
            return acc;
        }
        template <int B, class P, bool A0, int DI, int DJ, int DK, class C>
        __device__ __forceinline__ static double ev(const C& c) {
            double acc = 0.0;
            static_for<0, N2 - 1>([&](auto cz) {
                static_for<0, N1 - 1>([&](auto by) {
                    static_for<0, N0 - 1>([&](auto ax) {
                        constexpr int x = decltype(ax)::value, b = decltype(by)::value, z = decltype(cz)::value;
                        acc = P::add(acc, P::mul(c.a.s[K0 + x + N0 * (b + N1 * z)], E::template ev<B + 1, P, A0, DI + x - H0, DJ + b - H1, DK + z - H2>(c)));
                    });
                });
            });
            return acc;
        }
    };

    // ------------------------------------------------------------------------------------------- skeletons
    // compound assignment (BasicArithOp, Constants.hpp:53; FieldAssigner.hpp:48-80).  `op` is warp-uniform.
    template <class P>
    __device__ __forceinline__ double apply_op(int op, double oldv, double v) {
        switch (op) {
            case 1: return P::add(oldv, v);
            case 2: return P::sub(oldv, v);
            case 3: return P::mul(oldv, v);
            case 4: return P::div(oldv, v);
            default: return v;
        }
    }

    // K1: dst[i] (op)= expr.evalAt(i) over a box (FieldAssigner.hpp:48-80, rangeFor RangeFor.hpp:69-84).
    // Thread <-> one cell of the fastest axes; the block marches along the slowest axis (chunk `ch` per block) so
    // that the planes/rows shared by consecutive iterations are served by L1 and loop-invariant per-axis
    // coefficient loads are hoisted.  Coalescing: axis 0 is contiguous across the warp.
    // Many-operand expressions (the semi-implicit momentum operators: E::size > 48, the ones the register window is gated off for) are
    // latency-bound at 128 registers per thread -- one 512-thread block, 24 % of the warps an SM can hold (profiles/r2_momentum_ncu.md).
    // For them the kernel is compiled for three 256-thread blocks per SM (85 registers, ~1 KB of spill per thread): 9 % faster on the
    // C5 momentum solves.  Everything else keeps the 512-thread / 128-register form (WENO, compound ops: FP64- or HBM-bound).
    template <class E>
    inline constexpr bool assign_many_operands = E::size > 48;
    template <class E, class P, bool A0, int DIM>
    __device__ __forceinline__ void assign_body(const ExprArgs& a, const DstView& dst, const double* __restrict__ oldp, const LaunchRange& r, const int ch,
                                                const int op) {
        const int i = r.lo[0] + blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= r.hi[0]) return;
        if constexpr (DIM == 1) {
            double v = E::template eval<0, P, A0>(a, i, 0, 0);
            if (op != 0) v = apply_op<P>(op, oldp[i], v);
            dst.p[i] = v;
        } else if constexpr (DIM == 2) {
            const int j0 = r.lo[1] + blockIdx.y * ch;
            const int j1 = min(j0 + ch, r.hi[1]);
#pragma unroll 4
            for (int j = j0; j < j1; ++j) {
                double v = E::template eval<0, P, A0>(a, i, j, 0);
                const long long o = (long long) i + (long long) j * dst.s1;
                if (op != 0) v = apply_op<P>(op, oldp[o], v);
                dst.p[o] = v;
            }
        } else {
            const int j = r.lo[1] + blockIdx.y * blockDim.y + threadIdx.y;
            if (j >= r.hi[1]) return;
            const int k0 = r.lo[2] + blockIdx.z * ch;
            const int k1 = min(k0 + ch, r.hi[2]);
#pragma unroll 4
            for (int k = k0; k < k1; ++k) {
                double v = E::template eval<0, P, A0>(a, i, j, k);
                const long long o = (long long) i + (long long) j * dst.s1 + (long long) k * dst.s2;
                if (op != 0) v = apply_op<P>(op, oldp[o], v);
                dst.p[o] = v;
            }
        }
    }
    // two compilations of the same body: the default one (512 threads, the compiler's own register choice) and the many-operand one
    // (an explicit second launch-bounds argument also changes the register allocation of small kernels -- measured: + 18 % on the explicit
    // corrections of C5 -- so it is not applied to them)
    template <class E, class P, bool A0, int DIM>
    __global__ void __launch_bounds__(512) assign_kernel(const __grid_constant__ ExprArgs a, const DstView dst, const double* __restrict__ oldp,
                                                         const LaunchRange r, const int ch, const int op) {
        assign_body<E, P, A0, DIM>(a, dst, oldp, r, ch, op);
    }
    template <class E, class P, bool A0, int DIM>
    __global__ void __launch_bounds__(256, 3) assign_kernel_mo(const __grid_constant__ ExprArgs a, const DstView dst, const double* __restrict__ oldp,
                                                               const LaunchRange r, const int ch, const int op) {
        assign_body<E, P, A0, DIM>(a, dst, oldp, r, ch, op);
    }
    template <class E, class P, bool A0, int DIM, class G>
    inline void launch_assign_kernel(const G& g, cudaStream_t st, const ExprArgs& a, const DstView& dst, const double* oldp, const LaunchRange& r,
                                     int op) {
        if constexpr (assign_many_operands<E>) assign_kernel_mo<E, P, A0, DIM><<<g.grid, g.block, 0, st>>>(a, dst, oldp, r, g.ch, op);
        else
            assign_kernel<E, P, A0, DIM><<<g.grid, g.block, 0, st>>>(a, dst, oldp, r, g.ch, op);
    }

    // ------------------------------------------------------------------------------------------- register-window skeleton
    // K1, fast path.  Each thread owns CX consecutive cells along axis 0 and marches along the slowest axis.  For every
    // field slot the thread keeps a *register window*: the rows (cross offset dc, march offset dm) the expression taps
    // (known at compile time from E::taps), each CX + halo wide.  Per march step only the leading row of every column is
    // loaded -- 128-bit vector loads for the aligned core -- and the window is rotated in registers, so a 7-point
    // stencil issues 3 vector loads + 2 scalar halo loads per CX cells instead of 7 scalar loads per cell.

    template <class E, bool A0, int DIM>
    struct WinInfo {
        static constexpr int NS = A0 ? 1 : (E::nf > 0 ? E::nf : 1);
        static constexpr TapSet<NS> make() {
            TapSet<NS> ts;
            E::template taps<A0>(ts, tap_origin());
            return ts;
        }
        static constexpr TapSet<NS> ts = make();
        static constexpr bool overflow = ts.overflow;
        // (cross, march) -> (dj, dk): 3-D marches along k with cross axis j; 2-D marches along j
        __host__ __device__ static constexpr int dj_of(int dc, int dm) { return DIM == 3 ? dc : dm; }
        __host__ __device__ static constexpr int dk_of(int dc, int dm) { return DIM == 3 ? dm : 0; }
        __host__ __device__ static constexpr bool tap(int s, int di, int dc, int dm) {
            const int dj = dj_of(dc, dm), dk = dk_of(dc, dm);
            if (di < -WR || di > WR || dj < -WR || dj > WR || dk < -WR || dk > WR) return false;
            return ts.g[s].t[dk + WR][dj + WR][di + WR];
        }
        // column (s, dc): march extent and x extent over the whole column
        __host__ __device__ static constexpr bool col_used(int s, int dc) {
            for (int dm = -WR; dm <= WR; ++dm)
                for (int di = -WR; di <= WR; ++di)
                    if (tap(s, di, dc, dm)) return true;
            return false;
        }
        __host__ __device__ static constexpr int col_mlo(int s, int dc) {
            for (int dm = -WR; dm <= WR; ++dm)
                for (int di = -WR; di <= WR; ++di)
                    if (tap(s, di, dc, dm)) return dm;
            return 0;
        }
        __host__ __device__ static constexpr int col_mhi(int s, int dc) {
            for (int dm = WR; dm >= -WR; --dm)
                for (int di = -WR; di <= WR; ++di)
                    if (tap(s, di, dc, dm)) return dm;
            return 0;
        }
        __host__ __device__ static constexpr int col_xlo(int s, int dc) {
            for (int di = -WR; di <= 0; ++di)
                for (int dm = -WR; dm <= WR; ++dm)
                    if (tap(s, di, dc, dm)) return di;
            return 0;
        }
        __host__ __device__ static constexpr int col_xhi(int s, int dc) {
            for (int di = WR; di >= 0; --di)
                for (int dm = -WR; dm <= WR; ++dm)
                    if (tap(s, di, dc, dm)) return di;
            return 0;
        }
        __host__ __device__ static constexpr int bound(int which) {// 0 xlo 1 xhi 2 clo 3 chi 4 mlo 5 mhi over all slots
            int v = 0;
            for (int s = 0; s < NS; ++s)
                for (int dc = -WR; dc <= WR; ++dc) {
                    if (!col_used(s, dc)) continue;
                    if (which == 0) v = col_xlo(s, dc) < v ? col_xlo(s, dc) : v;
                    if (which == 1) v = col_xhi(s, dc) > v ? col_xhi(s, dc) : v;
                    if (which == 2) v = dc < v ? dc : v;
                    if (which == 3) v = dc > v ? dc : v;
                    if (which == 4) v = col_mlo(s, dc) < v ? col_mlo(s, dc) : v;
                    if (which == 5) v = col_mhi(s, dc) > v ? col_mhi(s, dc) : v;
                }
            return v;
        }
        static constexpr int XL = bound(0), XH = bound(1), CL = bound(2), CH = bound(3), ML = bound(4), MH = bound(5);
        // doubles a thread keeps live in its register window with CX = 2 cells per thread: every tapped column (slot, cross offset)
        // holds its march extent + the prefetched row, each row CX + x-halo wide
        __host__ __device__ static constexpr int live_doubles() {
            int n = 0;
            for (int s = 0; s < NS; ++s)
                for (int dc = (DIM == 3 ? -WR : 0); dc <= (DIM == 3 ? WR : 0); ++dc)// 2-D has no cross axis: one column per slot
                    if (col_used(s, dc)) n += (col_mhi(s, dc) - col_mlo(s, dc) + 2) * (2 + col_xhi(s, dc) - col_xlo(s, dc));
            return n;
        }
        // cross offsets only exist in 3-D.  Expressions whose window does not fit the register file (many operand fields: the
        // semi-implicit momentum operators of LidDriven3D.cpp keep ~15 fields x 3-5 rows) take the direct-global skeleton: the window
        // skeleton spilled ~2.9 KB of stack per thread on them and ran 3.5x slower (cuobjdump -res-usage; profiles/r2_c5.md)
        static constexpr bool ok = !overflow && (DIM == 3 || DIM == 2) && live_doubles() <= 56 && E::size <= 48;
    };

    template <class E, bool A0, int DIM, int CX, bool UNI_ = false>
    struct WinCtx {
        using WI = WinInfo<E, A0, DIM>;
        static constexpr bool UNI = UNI_;
        const ExprArgs& a;
        int i0, j, k;
        unsigned valign;
        double w[WI::NS][WI::MH - WI::ML + 2][WI::CH - WI::CL + 1][CX + WI::XH - WI::XL];// +1 march slot: prefetched row
        // mesh coefficients in registers: [array][offset + CR] relative to i0 (axis 0), the cross index and the march
        // index.  Every entry is loaded by load_coefs*(); entries no stencil node reads are dead code for the compiler.
        // (mesh arrays carry 16 doubles of slack on both sides so even a not-eliminated load stays in bounds)
        static constexpr int CR = WR + 2;
        double cf0[5][CX + 2 * CR], cfc[5][2 * CR + 1], cfm[5][2 * CR + 1];

        __device__ __forceinline__ WinCtx(const ExprArgs& a_, unsigned va) : a(a_), i0(0), j(0), k(0), valign(va) {}

        static constexpr int MAXIS = DIM - 1;               // march axis
        static constexpr int CAXIS = DIM == 3 ? 1 : -1;     // cross axis (3-D only)
        __device__ __forceinline__ static const double* arr_of(const AxisView& ax, int which) {
            return which == CF_X ? ax.x : (which == CF_DX ? ax.dx : (which == CF_RDX ? ax.rdx : (which == CF_RDXH ? ax.rdxh : ax.rdxc)));
        }
        template <int D, int ARR, int O>
        __device__ __forceinline__ double coef() const {
            if constexpr (UNI && ARR != CF_X) return a.ax[D].u[ARR];// uniform axis: constant-bank operand
            else if constexpr (UNI) return __ldg(a.ax[D].x + ((D == 0 ? i0 : (D == 1 ? j : k)) + O));
            else if constexpr (D == 0) return cf0[ARR][O + CR];
            else if constexpr (D == MAXIS) return cfm[ARR][O + CR];
            else return cfc[ARR][O + CR];
        }
        __device__ __forceinline__ void load_coefs_fixed() {// axis 0 and the cross axis: loop invariant
#pragma unroll
            for (int arr = 0; arr < 5; ++arr) {
                const double* p0 = arr_of(a.ax[0], arr) + i0;
#pragma unroll
                for (int o = -CR; o < CX + CR; ++o) cf0[arr][o + CR] = __ldg(p0 + o);
                if constexpr (CAXIS >= 0) {
                    const double* pc = arr_of(a.ax[CAXIS], arr) + j;
#pragma unroll
                    for (int o = -CR; o <= CR; ++o) cfc[arr][o + CR] = __ldg(pc + o);
                }
            }
        }
        __device__ __forceinline__ void load_coefs_march(int m) {// warp-uniform addresses
#pragma unroll
            for (int arr = 0; arr < 5; ++arr) {
                const double* pm = arr_of(a.ax[MAXIS], arr) + m;
#pragma unroll
                for (int o = -CR; o <= CR; ++o) cfm[arr][o + CR] = __ldg(pm + o);
            }
        }

        template <int S, int DI, int DJ, int DK>
        __device__ __forceinline__ double get() const {
            constexpr int dc = DIM == 3 ? DJ : 0, dm = DIM == 3 ? DK : DJ;
            static_assert(DIM == 3 || DK == 0, "2-D expression taps axis 2");
            return w[S][dm - WI::ML][dc - WI::CL][DI - WI::XL];
        }

        // load row (S, DC, DM) of the window at the current (i0, j, k)
        template <int S, int DC, int DM>
        __device__ __forceinline__ void load_row() {
            constexpr int xlo = WI::col_xlo(S, DC), xhi = WI::col_xhi(S, DC);
            constexpr int dj = WI::dj_of(DC, DM), dk = WI::dk_of(DC, DM);
            const FieldView& v = a.f[S];
            const double* rp = v.p + ((long long) i0 + (long long) (j + dj) * v.s1 + (long long) (k + dk) * v.s2);
            double* row = w[S][DM - WI::ML][DC - WI::CL];
            if ((valign >> S) & 1) {
#pragma unroll
                for (int c = 0; c < CX; c += 2) {
                    const double2 t = __ldg(reinterpret_cast<const double2*>(rp + c));
                    row[c - WI::XL] = t.x;
                    row[c + 1 - WI::XL] = t.y;
                }
            } else {
#pragma unroll
                for (int c = 0; c < CX; ++c) row[c - WI::XL] = __ldg(rp + c);
            }
#pragma unroll
            for (int x = xlo; x < 0; ++x) row[x - WI::XL] = __ldg(rp + x);
#pragma unroll
            for (int x = CX; x <= CX - 1 + xhi; ++x) row[x - WI::XL] = __ldg(rp + x);
        }

        // prologue: rows [mlo, mhi] of every column at the first march index; steady state: prefetch row mhi+1 (the
        // leading row of the NEXT march step) while the current step computes, then rotate the window down by one.
        __device__ __forceinline__ void load_prologue() {
            static_for<0, WI::NS - 1>([&](auto s) {
                static_for<WI::CL, WI::CH>([&](auto dc) {
                    constexpr int S = decltype(s)::value, DC = decltype(dc)::value;
                    if constexpr (WI::col_used(S, DC)) {
                        static_for<WI::col_mlo(S, DC), WI::col_mhi(S, DC)>([&](auto dm) { load_row<S, DC, decltype(dm)::value>(); });
                    }
                });
            });
        }
        __device__ __forceinline__ void prefetch_next() {
            static_for<0, WI::NS - 1>([&](auto s) {
                static_for<WI::CL, WI::CH>([&](auto dc) {
                    constexpr int S = decltype(s)::value, DC = decltype(dc)::value;
                    if constexpr (WI::col_used(S, DC)) load_row<S, DC, WI::col_mhi(S, DC) + 1>();
                });
            });
        }
        // L2 prefetch of the DRAM-unique stream: the leading row of the dc = 0 column of every slot, `pd` march steps
        // ahead (one request per 128-byte line: lanes whose tile starts a line).  Costs no registers, unlike a deeper
        // register pipeline, and turns the later vector load into an L2 hit.
        __device__ __forceinline__ void prefetch_l2(int pd) {
            static_for<0, WI::NS - 1>([&](auto s) {
                constexpr int S = decltype(s)::value;
                if constexpr (WI::col_used(S, 0)) {
                    constexpr int dm = WI::col_mhi(S, 0);
                    constexpr int dj = WI::dj_of(0, dm), dk = WI::dk_of(0, dm);
                    const FieldView& v = a.f[S];
                    const double* rp = v.p + ((long long) i0 + (long long) (j + dj + (DIM == 2 ? pd : 0)) * v.s1 +
                                              (long long) (k + dk + (DIM == 3 ? pd : 0)) * v.s2);
                    if ((reinterpret_cast<unsigned long long>(rp) & 127ull) < (unsigned long long) (CX * 8))
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(rp));
                }
            });
        }
        __device__ __forceinline__ void rotate() {
            static_for<0, WI::NS - 1>([&](auto s) {
                static_for<WI::CL, WI::CH>([&](auto dc) {
                    constexpr int S = decltype(s)::value, DC = decltype(dc)::value;
                    if constexpr (WI::col_used(S, DC)) {
                        constexpr int xlo = WI::col_xlo(S, DC), xhi = WI::col_xhi(S, DC);
                        static_for<WI::col_mlo(S, DC), WI::col_mhi(S, DC)>([&](auto dm) {
                            constexpr int DM = decltype(dm)::value;
#pragma unroll
                            for (int x = xlo; x <= CX - 1 + xhi; ++x)
                                w[S][DM - WI::ML][DC - WI::CL][x - WI::XL] = w[S][DM + 1 - WI::ML][DC - WI::CL][x - WI::XL];
                        });
                    }
                });
            });
        }
    };

    template <class E, class P, bool A0, int DIM, int CX, bool HASOP, bool UNI = false>
    __global__ void __launch_bounds__(128, OPF_WIN_MINBLOCKS) window_kernel(const __grid_constant__ ExprArgs a, const DstView dst,
                                                         const double* __restrict__ oldp, const LaunchRange r, const int ch,
                                                         const int op, const unsigned valign, const int dalign, const int pd) {
        using Ctx = WinCtx<E, A0, DIM, CX, UNI>;
        const int i0 = r.lo[0] + (blockIdx.x * blockDim.x + threadIdx.x) * CX;
        if (i0 >= r.hi[0]) return;
        Ctx c(a, valign);
        c.i0 = i0;
        int m0, m1;// march range
        if constexpr (DIM == 3) {
            c.j = r.lo[1] + blockIdx.y * blockDim.y + threadIdx.y;
            if (c.j >= r.hi[1]) return;
            m0 = r.lo[2] + blockIdx.z * ch;
            m1 = min(m0 + ch, r.hi[2]);
            c.k = m0;
        } else {
            m0 = r.lo[1] + blockIdx.y * ch;
            m1 = min(m0 + ch, r.hi[1]);
            c.j = m0;
            c.k = 0;
        }
        const bool full = i0 + CX <= r.hi[0];
        if constexpr (!UNI) c.load_coefs_fixed();
        c.load_prologue();
        for (int m = m0; m < m1; ++m) {
            if constexpr (DIM == 3) c.k = m;
            else
                c.j = m;
            if (m + 1 < m1) c.prefetch_next();// in flight during this step's arithmetic
            if (pd > 0 && m + pd < r.hi[DIM - 1]) c.prefetch_l2(pd);
            if constexpr (!UNI) c.load_coefs_march(m);
            double out[CX];
            static_for<0, CX - 1>([&](auto cc) { out[decltype(cc)::value] = E::template ev<0, P, A0, decltype(cc)::value, 0, 0>(c); });
            const long long o = (long long) i0 + (long long) c.j * dst.s1 + (long long) c.k * dst.s2;
            if constexpr (HASOP) {
#pragma unroll
                for (int x = 0; x < CX; ++x)
                    if (full || i0 + x < r.hi[0]) out[x] = apply_op<P>(op, oldp[o + x], out[x]);
            }
            if (full && dalign) {
#pragma unroll
                for (int x = 0; x < CX; x += 2) *reinterpret_cast<double2*>(dst.p + o + x) = make_double2(out[x], out[x + 1]);
            } else {
#pragma unroll
                for (int x = 0; x < CX; ++x)
                    if (i0 + x < r.hi[0]) dst.p[o + x] = out[x];
            }
            c.rotate();
        }
    }

    // ------------------------------------------------------------------------------------------- TMA tile skeleton (3-D)
    // K1, sm_100a path for 3-D fields.  A block owns a (BX*CX) x BY tile of the x-y plane and marches along z.  One elected
    // thread streams whole tile-plus-halo planes of every field slot into a shared-memory ring with TMA
    // (cp.async.bulk.tensor.3d, completion on an mbarrier); all threads evaluate the functor with every tap served from
    // shared memory at compile-time offsets.  Loads cost no registers and no LSU instructions, out-of-range halo cells
    // are zero-filled by the TMA unit, and the ring depth (planes in flight) hides HBM latency independent of occupancy.
    extern "C" int opf_internal_tensor_map(void* out128, const void* base, const unsigned long long* dims,
                                           const unsigned long long* strides_b, const unsigned* box, int rank);

    __device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned) __cvta_generic_to_shared(p); }
    __device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    }
    __device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    }
    __device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
        asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "OPF_WAIT_%=:\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                "@p bra OPF_DONE_%=;\n"
                "bra OPF_WAIT_%=;\n"
                "OPF_DONE_%=:\n"
                "}\n" ::"r"(smem_u32(bar)),
                "r"(parity)
                : "memory");
    }
    __device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* map, unsigned long long* bar, int x, int y, int z) {
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                             smem_u32(smem)),
                     "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
                     : "memory");
    }

    template <int NS>
    struct alignas(64) TmaMaps {
        CUtensorMap m[NS];
        int org[NS][3];
    };

    template <class E, bool A0, int CX, int BX, int BY>
    struct TmaGeom {
        using WI = WinInfo<E, A0, 3>;
        static constexpr int HXL = (-WI::XL + 1) / 2 * 2, HXH = (WI::XH + 1) / 2 * 2;// even halos keep pairs 16-byte aligned
        static constexpr int BW = BX * CX + HXL + HXH;                               // box width (doubles), even
        static constexpr int HYL = -WI::CL, HYH = WI::CH, BH = BY + HYL + HYH;
        static constexpr int NPL = WI::MH - WI::ML + 1;// planes a cell taps
        static constexpr int TILE_B = (BW * BH * 8 + 127) / 128 * 128;
        static constexpr int NS = WI::NS;
    };

    template <class E, bool A0, int CX, int BX, int BY, bool UNI_>
    struct TmaCtx {
        using G = TmaGeom<E, A0, CX, BX, BY>;
        using WI = typename G::WI;
        static constexpr bool UNI = UNI_;
        const ExprArgs& a;
        int i0, j, k;
        const double* pl[G::NS][G::NPL];// this thread's centre element (x = i0) in every tapped plane of every slot
        static constexpr int CR = WR + 2;
        double cf0[5][CX + 2 * CR], cfc[5][2 * CR + 1], cfm[5][2 * CR + 1];
        __device__ __forceinline__ TmaCtx(const ExprArgs& a_) : a(a_), i0(0), j(0), k(0) {}
        __device__ __forceinline__ static const double* arr_of(const AxisView& ax, int which) {
            return which == CF_X ? ax.x : (which == CF_DX ? ax.dx : (which == CF_RDX ? ax.rdx : (which == CF_RDXH ? ax.rdxh : ax.rdxc)));
        }
        template <int D, int ARR, int O>
        __device__ __forceinline__ double coef() const {
            if constexpr (UNI && ARR != CF_X) return a.ax[D].u[ARR];// constant-bank operand, no load, no register
            else if constexpr (UNI) return __ldg(a.ax[D].x + ((D == 0 ? i0 : (D == 1 ? j : k)) + O));
            else if constexpr (D == 0) return cf0[ARR][O + CR];
            else if constexpr (D == 2) return cfm[ARR][O + CR];
            else return cfc[ARR][O + CR];
        }
        __device__ __forceinline__ void load_coefs_fixed() {
#pragma unroll
            for (int arr = 0; arr < 5; ++arr) {
                const double* p0 = arr_of(a.ax[0], arr) + i0;
#pragma unroll
                for (int o = -CR; o < CX + CR; ++o) cf0[arr][o + CR] = __ldg(p0 + o);
                const double* pc = arr_of(a.ax[1], arr) + j;
#pragma unroll
                for (int o = -CR; o <= CR; ++o) cfc[arr][o + CR] = __ldg(pc + o);
            }
        }
        __device__ __forceinline__ void load_coefs_march(int m) {
#pragma unroll
            for (int arr = 0; arr < 5; ++arr) {
                const double* pm = arr_of(a.ax[2], arr) + m;
#pragma unroll
                for (int o = -CR; o <= CR; ++o) cfm[arr][o + CR] = __ldg(pm + o);
            }
        }
        // core[s][plane][row][c]: the thread's own CX cells of every tapped row, fetched once per march step with 128-bit
        // conflict-free shared loads; only x-halo taps (DI outside [0, CX)) read shared memory individually
        double core[G::NS][G::NPL][G::HYL + G::HYH + 1][CX];
        __host__ __device__ static constexpr bool row_used(int s, int dj, int dk) {
            for (int di = -WR; di <= WR; ++di)
                if (WI::tap(s, di, dj, dk)) return true;
            return false;
        }
        // FIRST: every tapped row comes from shared memory.  Later march steps: a row (dj, dk) whose successor (dj, dk+1)
        // was held last step is that successor's registers (the window slides by one plane) -- only rows without a held
        // successor (the newest plane, and rows tapped in a single plane) are read from shared memory again.
        template <bool FIRST>
        __device__ __forceinline__ void load_cores() {
            static_for<0, G::NS - 1>([&](auto s) {
                static_for<WI::ML, WI::MH>([&](auto dk) {// ascending: [dk+1] still holds last step's values when copied
                    static_for<WI::CL, WI::CH>([&](auto dj) {
                        constexpr int S = decltype(s)::value, DK = decltype(dk)::value, DJ = decltype(dj)::value;
                        if constexpr (row_used(S, DJ, DK)) {
                            if constexpr (!FIRST && DK < WI::MH && row_used(S, DJ, DK + 1)) {
#pragma unroll
                                for (int c = 0; c < CX; ++c) core[S][DK - WI::ML][DJ + G::HYL][c] = core[S][DK + 1 - WI::ML][DJ + G::HYL][c];
                            } else {
#pragma unroll
                                for (int c = 0; c < CX; c += 2) {
                                    const double2 t = *reinterpret_cast<const double2*>(pl[S][DK - WI::ML] + DJ * G::BW + c);
                                    core[S][DK - WI::ML][DJ + G::HYL][c] = t.x;
                                    core[S][DK - WI::ML][DJ + G::HYL][c + 1] = t.y;
                                }
                            }
                        }
                    });
                });
            });
        }
        template <int S, int DI, int DJ, int DK>
        __device__ __forceinline__ double get() const {
            if constexpr (DI >= 0 && DI < CX) return core[S][DK - WI::ML][DJ + G::HYL][DI];
            else return pl[S][DK - WI::ML][DJ * G::BW + DI];
        }
    };

    template <class E, class P, bool A0, int CX, int BX, int BY, int STAGES, bool HASOP, bool UNI>
    __global__ void __launch_bounds__(BX* BY) tma_kernel(const __grid_constant__ ExprArgs a,
                                                         const __grid_constant__ TmaMaps<TmaGeom<E, A0, CX, BX, BY>::NS> maps, const DstView dst,
                                                         const double* __restrict__ oldp, const LaunchRange r, const int ch, const int op,
                                                         const int dalign) {
        using G = TmaGeom<E, A0, CX, BX, BY>;
        using WI = typename G::WI;
        using Ctx = TmaCtx<E, A0, CX, BX, BY, UNI>;
        constexpr int NS = G::NS, NPL = G::NPL;
        static_assert(STAGES > NPL, "ring must hold the tapped planes plus at least one plane in flight");
        extern __shared__ __align__(128) unsigned char opf_tma_smem[];
        unsigned long long* full = reinterpret_cast<unsigned long long*>(opf_tma_smem);// [STAGES]
        unsigned char* ring = opf_tma_smem + 128;                                       // [STAGES][NS] tiles of TILE_B bytes
        const int tid = threadIdx.y * BX + threadIdx.x;
        const int x0 = r.lo[0] + blockIdx.x * (BX * CX), y0 = r.lo[1] + blockIdx.y * BY;
        const int m0 = r.lo[2] + blockIdx.z * ch, m1 = min(m0 + ch, r.hi[2]);
        const int pfirst = m0 + WI::ML, plast = m1 - 1 + WI::MH;// planes this block touches
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncthreads();
        auto issue = [&](int plane, int stage) {// elected thread: one box per slot, all completing on full[stage]
            mbar_expect_tx(&full[stage], (unsigned) (NS * G::BW * G::BH * 8));
#pragma unroll
            for (int s = 0; s < NS; ++s)
                tma_load_3d(ring + (size_t) (stage * NS + s) * G::TILE_B, &maps.m[s], &full[stage], x0 - G::HXL - maps.org[s][0],
                            y0 - G::HYL - maps.org[s][1], plane - maps.org[s][2]);
        };
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < STAGES; ++s)
                if (pfirst + s <= plast) issue(pfirst + s, s);
        }
        Ctx c(a);
        c.i0 = x0 + threadIdx.x * CX;
        c.j = y0 + threadIdx.y;
        const bool active = c.i0 < r.hi[0] && c.j < r.hi[1];
        const bool full_tile = c.i0 + CX <= r.hi[0];
        if constexpr (!UNI)
            if (active) c.load_coefs_fixed();
        // byte offset of this thread's centre element inside a tile
        const int toff = ((threadIdx.y + G::HYL) * G::BW + G::HXL + threadIdx.x * CX) * 8;
        // wait for the planes the first step taps except the newest one (waited inside the loop)
        int s_lo = 0;             // stage holding plane m + ML
        unsigned phase_bits = 0u; // bit s: parity to wait for on full[s]
#pragma unroll
        for (int q = 0; q < NPL - 1; ++q) {
            mbar_wait(&full[q], 0u);
        }
        for (int q = 0; q < NPL - 1; ++q) phase_bits ^= 1u << q;
        for (int m = m0; m < m1; ++m) {
            c.k = m;
            // newest tapped plane m + MH lives in stage (s_lo + NPL - 1) % STAGES
            int s_new = s_lo + NPL - 1;
            if (s_new >= STAGES) s_new -= STAGES;
            mbar_wait(&full[s_new], (phase_bits >> s_new) & 1u);
            phase_bits ^= 1u << s_new;
            if (active) {
#pragma unroll
                for (int q = 0; q < NPL; ++q) {
                    int st = s_lo + q;
                    if (st >= STAGES) st -= STAGES;
#pragma unroll
                    for (int s = 0; s < NS; ++s)
                        c.pl[s][q] = reinterpret_cast<const double*>(ring + (size_t) (st * NS + s) * G::TILE_B + toff);
                }
                if (m == m0) c.template load_cores<true>();
                else
                    c.template load_cores<false>();
                if constexpr (!UNI) c.load_coefs_march(m);
                double out[CX];
                static_for<0, CX - 1>([&](auto cc) { out[decltype(cc)::value] = E::template ev<0, P, A0, decltype(cc)::value, 0, 0>(c); });
                const long long o = (long long) c.i0 + (long long) c.j * dst.s1 + (long long) m * dst.s2;
                if constexpr (HASOP) {
#pragma unroll
                    for (int x = 0; x < CX; ++x)
                        if (full_tile || c.i0 + x < r.hi[0]) out[x] = apply_op<P>(op, oldp[o + x], out[x]);
                }
                if (full_tile && dalign) {
#pragma unroll
                    for (int x = 0; x < CX; x += 2) *reinterpret_cast<double2*>(dst.p + o + x) = make_double2(out[x], out[x + 1]);
                } else {
#pragma unroll
                    for (int x = 0; x < CX; ++x)
                        if (c.i0 + x < r.hi[0]) dst.p[o + x] = out[x];
                }
            }
            __syncthreads();// every thread is done with plane m + ML: its stage can be refilled
            if (tid == 0) {
                const int pnext = m + WI::ML + STAGES;
                if (pnext <= plast) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    issue(pnext, s_lo);
                }
            }
            s_lo = s_lo + 1 == STAGES ? 0 : s_lo + 1;
        }
    }

    // K7: rangeReduce (RangeFor.hpp:87-121).  Warp-shuffle tree + one partial per block; a second tiny launch folds
    // the partials in a fixed order, so results are run-to-run deterministic.  rop: opf_reduce_op (warp-uniform).
    __device__ __forceinline__ double red_identity(int rop) { return rop == 1 ? -INFINITY : (rop == 2 ? INFINITY : 0.0); }
    __device__ __forceinline__ double red_lift(int rop, double v) { return rop == 3 ? fabs(v) : (rop == 4 ? v * v : v); }
    __device__ __forceinline__ double red_comb(int rop, double x, double y) {
        return (rop == 1 || rop == 3) ? fmax(x, y) : (rop == 2 ? fmin(x, y) : x + y);
    }
    __device__ __forceinline__ double block_reduce(int rop, double v) {
        __shared__ double sh[32];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = red_comb(rop, v, __shfl_xor_sync(0xffffffffu, v, o));
        const int tid = threadIdx.x, nth = blockDim.x;
        const int w = tid >> 5, l = tid & 31;
        if (l == 0) sh[w] = v;
        __syncthreads();
        if (w == 0) {
            v = l < ((nth + 31) >> 5) ? sh[l] : red_identity(rop);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v = red_comb(rop, v, __shfl_xor_sync(0xffffffffu, v, o));
        }
        return v;// valid in thread 0
    }

    // work item = (row j,k ; 2048-wide segment of axis 0); blocks stride over items, threads over the segment (coalesced)
    constexpr int RED_SEG = 2048;
    template <class E, class P, bool A0>
    __global__ void __launch_bounds__(256) reduce_kernel(const __grid_constant__ ExprArgs a, const LaunchRange r,
                                                         double* __restrict__ partials, const int rop) {
        const long long n0 = r.hi[0] - r.lo[0], n1 = r.hi[1] - r.lo[1], n2 = r.hi[2] - r.lo[2];
        const long long nseg = (n0 + RED_SEG - 1) / RED_SEG;
        const long long items = n1 * n2 * nseg;
        double acc = red_identity(rop);
        for (long long it = blockIdx.x; it < items; it += gridDim.x) {
            const long long seg = it % nseg, row = it / nseg;
            const int j = r.lo[1] + (int) (row % n1), k = r.lo[2] + (int) (row / n1);
            const long long b = seg * RED_SEG, e = min(b + (long long) RED_SEG, n0);
            // four independent evaluations per trip: enough loads in flight to stream at HBM rate with 4 blocks per SM
            long long ii = b + threadIdx.x;
            const long long st = blockDim.x;
            for (; ii + 3 * st < e; ii += 4 * st) {
                const double v0 = red_lift(rop, E::template eval<0, P, A0>(a, r.lo[0] + (int) ii, j, k));
                const double v1 = red_lift(rop, E::template eval<0, P, A0>(a, r.lo[0] + (int) (ii + st), j, k));
                const double v2 = red_lift(rop, E::template eval<0, P, A0>(a, r.lo[0] + (int) (ii + 2 * st), j, k));
                const double v3 = red_lift(rop, E::template eval<0, P, A0>(a, r.lo[0] + (int) (ii + 3 * st), j, k));
                acc = red_comb(rop, red_comb(rop, red_comb(rop, red_comb(rop, acc, v0), v1), v2), v3);
            }
            for (; ii < e; ii += st) acc = red_comb(rop, acc, red_lift(rop, E::template eval<0, P, A0>(a, r.lo[0] + (int) ii, j, k)));
        }
        acc = block_reduce(rop, acc);
        if (threadIdx.x == 0) partials[blockIdx.x] = acc;
    }

    // partials are already lifted (|.|, .^2): the final fold only combines
    template <int U = 0>
    __global__ void __launch_bounds__(256) reduce_final_kernel(const double* __restrict__ partials, int n,
                                                                      double* __restrict__ out, const int rop) {
        const int crop = rop == 3 ? 1 : (rop == 4 ? 0 : rop);
        double acc = red_identity(crop);
        for (int i = threadIdx.x; i < n; i += blockDim.x) acc = red_comb(crop, acc, partials[i]);
        acc = block_reduce(crop, acc);
        if (threadIdx.x == 0) out[0] = acc;
    }

    // ------------------------------------------------------------------------------------------- launcher
    // One launcher per expression type: dispatches (mode, alias0, dim) to the kernel instantiations.  Its address is
    // what opf_expr_register() stores; the engine calls it with the blobs it prepared (ExprArgs, LaunchInfo).
    struct LaunchGeom {
        dim3 grid, block;
        int ch;
    };
    inline int pick_tx(int n0) {
        // widest x-tile whose last block wastes < half a tile
        int tx = 128;
        while (tx > 32 && ((n0 + tx - 1) / tx) * tx - n0 > tx / 2) tx >>= 1;
        return tx;
    }
    inline LaunchGeom assign_geometry(const LaunchInfo& li, int max_threads = 512) {
        LaunchGeom g;
        const int n0 = li.r.hi[0] - li.r.lo[0], n1 = li.r.hi[1] - li.r.lo[1], n2 = li.r.hi[2] - li.r.lo[2];
        if (n0 <= 0 || n1 <= 0 || n2 <= 0) {
            g.grid = dim3(0, 0, 0);
            g.block = dim3(1, 1, 1);
            g.ch = 1;
            return g;
        }
        if (li.dim == 1) {
            g.block = dim3(256, 1, 1);
            g.grid = dim3((n0 + 255) / 256, 1, 1);
            g.ch = 1;
        } else if (li.dim == 2) {
            const int tx = n0 >= 1024 ? 256 : pick_tx(n0);
            g.block = dim3(tx, 1, 1);
            g.ch = 16;
            g.grid = dim3((n0 + tx - 1) / tx, (n1 + g.ch - 1) / g.ch, 1);
        } else {
            // tunables for sweeps (read once): OPF_TX / OPF_TY / OPF_CH
            static const int etx = getenv("OPF_TX") ? atoi(getenv("OPF_TX")) : 0;
            static const int ety = getenv("OPF_TY") ? atoi(getenv("OPF_TY")) : 0;
            static const int ech = getenv("OPF_CH") ? atoi(getenv("OPF_CH")) : 0;
            const int tx = etx > 0 ? etx : pick_tx(n0);
            int ty = ety > 0 ? ety : 4;
            while (ty > 1 && tx * ty > max_threads) ty >>= 1;
            g.block = dim3(tx, ty, 1);
            g.ch = ech > 0 ? ech : 32;
            g.grid = dim3((n0 + tx - 1) / tx, (n1 + ty - 1) / ty, (n2 + g.ch - 1) / g.ch);
        }
        return g;
    }

    template <class E, class P, bool A0, int DIM>
    int launch_window(const ExprArgs& a, const LaunchInfo& li, cudaStream_t st) {
        constexpr int CX = 2;
        static const int ety = getenv("OPF_WTY") ? atoi(getenv("OPF_WTY")) : 0;
        static const int etx = getenv("OPF_WTX") ? atoi(getenv("OPF_WTX")) : 0;
        static const int ech = getenv("OPF_WCH") ? atoi(getenv("OPF_WCH")) : 0;
        const int n0 = li.r.hi[0] - li.r.lo[0], n1 = li.r.hi[1] - li.r.lo[1], n2 = li.r.hi[2] - li.r.lo[2];
        const int nt = (n0 + CX - 1) / CX;// thread tiles along x
        int tx = etx > 0 ? etx : 32;
        dim3 block, grid;
        int ch;
        // march chunk: 16 steps per thread (measured on FTCS2D 4097^2: 51 us at 16 vs 61 us at 64 vs 92 us at 128 -- shorter
        // chunks give more, shorter blocks and a better balance over the 148 SMs than the 2-row prologue costs); small boxes
        // (coarse multigrid levels) get even shorter chunks so that the grid still covers the SMs
        auto pick_ch = [&](long long blocks_xy, int nmarch) {
            int c = 16;
            while (c > 4 && blocks_xy * ((nmarch + c - 1) / c) < 4 * 148) c >>= 1;
            return c;
        };
        if (DIM == 3) {
            const int ty = ety > 0 ? ety : 128 / tx;
            block = dim3(tx, ty, 1);
            const long long bxy = (long long) ((nt + tx - 1) / tx) * ((n1 + ty - 1) / ty);
            ch = ech > 0 ? ech : pick_ch(bxy, n2);
            grid = dim3((nt + tx - 1) / tx, (n1 + ty - 1) / ty, (n2 + ch - 1) / ch);
        } else {
            tx = etx > 0 ? etx : (nt >= 128 ? 128 : (nt >= 64 ? 64 : 32));// <= 128 (launch bounds)
            ch = ech > 0 ? ech : pick_ch((nt + tx - 1) / tx, n1);
            block = dim3(tx, 1, 1);
            grid = dim3((nt + tx - 1) / tx, (n1 + ch - 1) / ch, 1);
        }
        static const int pd = getenv("OPF_WPD") ? atoi(getenv("OPF_WPD")) : 8;
        auto go = [&](auto kern) {
            kern<<<grid, block, 0, st>>>(a, li.dst, li.old, li.r, ch, li.op, li.valign, li.dalign, pd);
            opf_internal_note_kernel("opf::window_kernel");
            return (int) cudaGetLastError();
        };
        if (li.uniform) {// single-spacing axes: coefficients from the constant bank
            if (li.op != 0) return go(window_kernel<E, P, A0, DIM, CX, true, true>);
            return go(window_kernel<E, P, A0, DIM, CX, false, true>);
        }
        if (li.op != 0) return go(window_kernel<E, P, A0, DIM, CX, true, false>);
        return go(window_kernel<E, P, A0, DIM, CX, false, false>);
    }

    // tile shape of the TMA skeleton: each thread owns OPF_TMA_CX consecutive cells of a row, a block is (128 / OPF_TMA_CX) x OPF_TMA_BY threads
#ifndef OPF_TMA_CX
#define OPF_TMA_CX 2
#endif
#ifndef OPF_TMA_BY
#define OPF_TMA_BY 4
#endif
    template <class E, bool A0, int CX, int BY>
    struct TmaFits {// ring of (planes tapped + 3) tile-plus-halo planes per slot must fit the shared-memory budget
        using G = TmaGeom<E, A0, CX, 128 / CX, BY>;
        static constexpr bool value = 128 + (G::NPL + 3) * G::NS * G::TILE_B <= 200 * 1024;
    };
    template <class E, class P, bool A0, int BY, int CX>
    int launch_tma(const ExprArgs& a, const LaunchInfo& li, cudaStream_t st) {
        constexpr int BX = 128 / CX;// 128-cell rows: the tile-plus-halo box stays under TMA's 256-element box limit
        using G = TmaGeom<E, A0, CX, BX, BY>;
        constexpr int STAGES = G::NPL + 3;
        static const int ech = getenv("OPF_TCH") ? atoi(getenv("OPF_TCH")) : 0;
        TmaMaps<G::NS> maps;
        for (int s = 0; s < G::NS; ++s) {
            const auto& t = li.tma[s];
            const unsigned box[3] = {(unsigned) G::BW, (unsigned) G::BH, 1u};
            if (opf_internal_tensor_map(&maps.m[s], t.base, t.dim, t.stride_b, box, 3) != 0) return -3;
            for (int d = 0; d < 3; ++d) maps.org[s][d] = t.org[d];
        }
        const int n0 = li.r.hi[0] - li.r.lo[0], n1 = li.r.hi[1] - li.r.lo[1], n2 = li.r.hi[2] - li.r.lo[2];
        // march chunk per block: short chunks (16 planes) measured best on B200 (513^3: 0.371 ms vs 0.404 ms at 64) -- more,
        // shorter blocks balance the 148 SMs better than the 2-plane ring prologue costs
        const int ch = ech > 0 ? ech : 16;
        dim3 block(BX, BY, 1), grid((n0 + BX * CX - 1) / (BX * CX), (n1 + BY - 1) / BY, (n2 + ch - 1) / ch);
        const size_t smem = 128 + (size_t) STAGES * G::NS * G::TILE_B;
        auto go = [&](auto kern) {
            if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
            kern<<<grid, block, smem, st>>>(a, maps, li.dst, li.old, li.r, ch, li.op, li.dalign);
            opf_internal_note_kernel("opf::tma_kernel");
            return (int) cudaGetLastError();
        };
        if (li.uniform) {
            if (li.op != 0) return go(tma_kernel<E, P, A0, CX, BX, BY, STAGES, true, true>);
            return go(tma_kernel<E, P, A0, CX, BX, BY, STAGES, false, true>);
        }
        if (li.op != 0) return go(tma_kernel<E, P, A0, CX, BX, BY, STAGES, true, false>);
        return go(tma_kernel<E, P, A0, CX, BX, BY, STAGES, false, false>);
    }

    // DIMS: bit (d-1) set <=> fields of dimension d can reach this launcher (the library's builtins serve all three; the
    // front-end knows its field's dimension at compile time and instantiates only that one)
    template <class E, class P, bool A0, int DIMS>
    int launch_assign(const ExprArgs& a, const LaunchInfo& li, cudaStream_t st) {
        const LaunchGeom g = assign_geometry(li, assign_many_operands<E> ? 256 : 512);
        if (g.grid.x == 0) return 0;
        if constexpr (E::maxaxis < 1 && (DIMS & 1))
            if (li.dim == 1) {
                launch_assign_kernel<E, P, A0, 1>(g, st, a, li.dst, li.old, li.r, li.op);
                opf_internal_note_kernel("opf::assign_kernel");
                return (int) cudaGetLastError();
            }
        if constexpr (E::maxaxis < 2 && (DIMS & 2))
            if (li.dim == 2) {
                if constexpr (WinInfo<E, A0, 2>::ok && E::nf > 0)
                    if (li.window) return launch_window<E, P, A0, 2>(a, li, st);
                launch_assign_kernel<E, P, A0, 2>(g, st, a, li.dst, li.old, li.r, li.op);
                opf_internal_note_kernel("opf::assign_kernel");
                return (int) cudaGetLastError();
            }
        if constexpr ((DIMS & 4) != 0)
        if (li.dim == 3) {
            if constexpr (WinInfo<E, A0, 3>::ok && E::nf > 0) {
                // TMA tile skeleton: footprint must fit the ring/box budget (<= 200 KB of shared memory)
                if (opf_internal_opt(OPF_OPT_TMA) && li.window && li.tma_ok && (li.r.hi[0] - li.r.lo[0]) >= 64) {
#ifdef OPF_TMA_SWEEP
                    // tuning build: tile shape selectable at run time (OPF_TBY in {4,8}, OPF_TCX in {2,4})
                    static const int tby = getenv("OPF_TBY") ? atoi(getenv("OPF_TBY")) : OPF_TMA_BY;
                    static const int tcx = getenv("OPF_TCX") ? atoi(getenv("OPF_TCX")) : OPF_TMA_CX;
                    if constexpr (TmaFits<E, A0, 4, 8>::value)
                        if (tby == 8 && tcx == 4) return launch_tma<E, P, A0, 8, 4>(a, li, st);
                    if constexpr (TmaFits<E, A0, 2, 8>::value)
                        if (tby == 8 && tcx == 2) return launch_tma<E, P, A0, 8, 2>(a, li, st);
                    if constexpr (TmaFits<E, A0, 4, 4>::value)
                        if (tby == 4 && tcx == 4) return launch_tma<E, P, A0, 4, 4>(a, li, st);
                    if constexpr (TmaFits<E, A0, 2, 4>::value)
                        if (tby == 4 && tcx == 2) return launch_tma<E, P, A0, 4, 2>(a, li, st);
#endif
                    if constexpr (TmaFits<E, A0, OPF_TMA_CX, OPF_TMA_BY>::value) return launch_tma<E, P, A0, OPF_TMA_BY, OPF_TMA_CX>(a, li, st);
                    else if constexpr (TmaFits<E, A0, 2, 4>::value) return launch_tma<E, P, A0, 4, 2>(a, li, st);
                }
                if (li.window) return launch_window<E, P, A0, 3>(a, li, st);
            }
            launch_assign_kernel<E, P, A0, 3>(g, st, a, li.dst, li.old, li.r, li.op);
            opf_internal_note_kernel("opf::assign_kernel");
            return (int) cudaGetLastError();
        }
        return -2;// expression uses an axis the field does not have
    }
    // direct-global skeleton only (Stencil arithmetic: a verification mode, not a fast path)
    template <class E, class P, bool A0, int DIMS>
    int launch_assign_plain(const ExprArgs& a, const LaunchInfo& li, cudaStream_t st) {
        const LaunchGeom g = assign_geometry(li, assign_many_operands<E> ? 256 : 512);
        if (g.grid.x == 0) return 0;
        opf_internal_note_kernel("opf::assign_kernel");
        if constexpr (E::maxaxis < 1 && (DIMS & 1))
            if (li.dim == 1) {
                launch_assign_kernel<E, P, A0, 1>(g, st, a, li.dst, li.old, li.r, li.op);
                return (int) cudaGetLastError();
            }
        if constexpr (E::maxaxis < 2 && (DIMS & 2))
            if (li.dim == 2) {
                launch_assign_kernel<E, P, A0, 2>(g, st, a, li.dst, li.old, li.r, li.op);
                return (int) cudaGetLastError();
            }
        if constexpr ((DIMS & 4) != 0)
            if (li.dim == 3) {
                launch_assign_kernel<E, P, A0, 3>(g, st, a, li.dst, li.old, li.r, li.op);
                return (int) cudaGetLastError();
            }
        return -2;
    }
    template <class E, class P, bool A0>
    int launch_reduce(const ExprArgs& a, const LaunchInfo& li, cudaStream_t st) {
        const int nb = li.n_partials;
        reduce_kernel<E, P, A0><<<nb, 256, 0, st>>>(a, li.r, li.partials, li.rop);
        reduce_final_kernel<0><<<1, 256, 0, st>>>(li.partials, nb, li.result ? li.result : li.partials + nb, li.rop);
        return (int) cudaGetLastError();
    }

    // LaunchInfo.rop < 0 -> assignment; >= 0 -> reduction
    template <class E, int DIMS = 7>
    int launcher(const void* args_blob, const void* launch_blob, void* stream) {
        const ExprArgs& a = *static_cast<const ExprArgs*>(args_blob);
        const LaunchInfo& li = *static_cast<const LaunchInfo*>(launch_blob);
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        constexpr bool can_alias = E::nf > 1;
        if (li.rop >= 0) {
            if constexpr (can_alias)
                if (li.alias0) return li.mode != 1 ? launch_reduce<E, Exact, true>(a, li, st) : launch_reduce<E, Fast, true>(a, li, st);
            return li.mode != 1 ? launch_reduce<E, Exact, false>(a, li, st) : launch_reduce<E, Fast, false>(a, li, st);
        }
        if (li.mode == 2) {// OPF_MODE_STENCIL
            if constexpr (can_alias)
                if (li.alias0) return launch_assign_plain<E, Stencil, true, DIMS>(a, li, st);
            return launch_assign_plain<E, Stencil, false, DIMS>(a, li, st);
        }
        if constexpr (can_alias)
            if (li.alias0) return li.mode == 0 ? launch_assign<E, Exact, true, DIMS>(a, li, st) : launch_assign<E, Fast, true, DIMS>(a, li, st);
        return li.mode == 0 ? launch_assign<E, Exact, false, DIMS>(a, li, st) : launch_assign<E, Fast, false, DIMS>(a, li, st);
    }

}// namespace opf
