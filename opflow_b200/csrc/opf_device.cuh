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
//            (<= 1e-12 relative vs the reference).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <math.h>
#include <cstdlib>

namespace opf {

    constexpr int MAX_FIELDS = 16;
    constexpr int MAX_SCALARS = 16;
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
    struct AxisView {
        const double *x, *dx, *rdx, *rdxh, *rdxc;
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
        // reduction launches
        int rop;
        double* partials;// >= grid blocks
        int n_partials;
    };

    // ------------------------------------------------------------------------------------------- policies
    struct Exact {
        static constexpr bool fast = false;
        __device__ __forceinline__ static double add(double a, double b) { return __dadd_rn(a, b); }
        __device__ __forceinline__ static double sub(double a, double b) { return __dsub_rn(a, b); }
        __device__ __forceinline__ static double mul(double a, double b) { return __dmul_rn(a, b); }
        __device__ __forceinline__ static double div(double a, double b) { return __ddiv_rn(a, b); }
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
#define OPF_SHIFT(D, n) i + ((D) == 0 ? (n) : 0), j + ((D) == 1 ? (n) : 0), k + ((D) == 2 ? (n) : 0)

    // ------------------------------------------------------------------------------------------- leaves
    // CartesianField::evalAtImpl_final (CartesianField.hpp:775-780)
    template <int K>
    struct F {
        static constexpr int size = 1, maxaxis = -1, nf = K + 1;
        template <int B, class P, bool A0>
        __device__ __forceinline__ static double eval(const ExprArgs& a, int i, int j, int k) {
            const FieldView& v = a.f[A0 ? 0 : K];
            return __ldg(v.p + ((long long) i + (long long) j * v.s1 + (long long) k * v.s2));
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
    };

    // ------------------------------------------------------------------------------------------- point-wise
    // BinOpDefMacros.hpp.in:15-17 (operand order preserved), AMDS.hpp:34-91, MinMax.hpp:51-52, Compare.hpp, Boolean.hpp
#define OPF_BINOP(Name, EXPR)                                                                                          \
    template <class L, class R>                                                                                        \
    struct Name {                                                                                                      \
        static constexpr int size = 1 + L::size + R::size;                                                             \
        static constexpr int maxaxis = L::maxaxis > R::maxaxis ? L::maxaxis : R::maxaxis;                              \
        static constexpr int nf = L::nf > R::nf ? L::nf : R::nf;                                                       \
        template <int B, class P, bool A0>                                                                             \
        __device__ __forceinline__ static double eval(const ExprArgs& a, int i, int j, int k) {                        \
            const double x = L::template eval<B + 1, P, A0>(a, i, j, k);                                               \
            const double y = R::template eval<B + 1 + L::size, P, A0>(a, i, j, k);                                     \
            return EXPR;                                                                                               \
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
#undef OPF_BINOP

#define OPF_UNIOP(Name, EXPR)                                                                                          \
    template <class E>                                                                                                 \
    struct Name {                                                                                                      \
        static constexpr int size = 1 + E::size, maxaxis = E::maxaxis, nf = E::nf;                                     \
        template <int B, class P, bool A0>                                                                             \
        __device__ __forceinline__ static double eval(const ExprArgs& a, int i, int j, int k) {                        \
            const double x = E::template eval<B + 1, P, A0>(a, i, j, k);                                               \
            return EXPR;                                                                                               \
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
#undef OPF_UNIOP

    // CondOp::eval (Conditional.hpp:37-40)
    template <class C, class A, class Bb>
    struct Cond {
        static constexpr int size = 1 + C::size + A::size + Bb::size;
        static constexpr int maxaxis = (C::maxaxis > A::maxaxis ? C::maxaxis : A::maxaxis) > Bb::maxaxis
                                               ? (C::maxaxis > A::maxaxis ? C::maxaxis : A::maxaxis)
                                               : Bb::maxaxis;
        static constexpr int nf = (C::nf > A::nf ? C::nf : A::nf) > Bb::nf ? (C::nf > A::nf ? C::nf : A::nf) : Bb::nf;
        template <int B, class P, bool A0>
        __device__ __forceinline__ static double eval(const ExprArgs& a, int i, int j, int k) {
            const double c = C::template eval<B + 1, P, A0>(a, i, j, k);
            return c != 0.0 ? A::template eval<B + 1 + C::size, P, A0>(a, i, j, k)
                            : Bb::template eval<B + 1 + C::size + A::size, P, A0>(a, i, j, k);
        }
    };

    // ------------------------------------------------------------------------------------------- stencils
    // D2SecondOrderCentered<d>::eval (D2SecondOrderCentered.hpp:161-171)
    template <int D, class E>
    struct D2C {
        static constexpr int size = 1 + E::size, maxaxis = (D > E::maxaxis ? D : E::maxaxis), nf = E::nf;
        template <int B, class P, bool A0>
        __device__ __forceinline__ static double eval(const ExprArgs& a, int i, int j, int k) {
            const double l = E::template eval<B + 1, P, A0>(a, OPF_SHIFT(D, -1));
            const double c = E::template eval<B + 1, P, A0>(a, i, j, k);
            const double r = E::template eval<B + 1, P, A0>(a, OPF_SHIFT(D, 1));
            const bool center = (a.loc[B] >> D) & 1;
            const int q = axis_of<D>(i, j, k);
            const AxisView& ax = a.ax[D];
            if constexpr (P::fast) {
                if (!center) return ((r - c) * __ldg(ax.rdx + q) - (c - l) * __ldg(ax.rdx + q - 1)) * __ldg(ax.rdxh + q);
                return ((r - c) * __ldg(ax.rdxh + q + 1) - (c - l) * __ldg(ax.rdxh + q)) * __ldg(ax.rdxc + q);
            } else {
                double dxl, dxr;
                if (!center) {
                    dxl = __ldg(ax.dx + q - 1);
                    dxr = __ldg(ax.dx + q);
                } else {
                    dxl = P::mul(P::add(__ldg(ax.dx + q - 1), __ldg(ax.dx + q)), 0.5);
                    dxr = P::mul(P::add(__ldg(ax.dx + q), __ldg(ax.dx + q + 1)), 0.5);
                }
                const double dxc = P::mul(P::add(dxl, dxr), 0.5);
                return P::div(P::sub(P::div(P::sub(r, c), dxr), P::div(P::sub(c, l), dxl)), dxc);
            }
        }
    };

    // D1FirstOrderCentered<d>::eval (D1FirstOrderCentered.hpp:31-36); result loc flipped in prepare (:46)
    template <int D, class E>
    struct D1C {
        static constexpr int size = 1 + E::size, maxaxis = (D > E::maxaxis ? D : E::maxaxis), nf = E::nf;
        template <int B, class P, bool A0>
        __device__ __forceinline__ static double eval(const ExprArgs& a, int i, int j, int k) {
            const bool center = (a.loc[B] >> D) & 1;
            const int q = axis_of<D>(i, j, k);
            const AxisView& ax = a.ax[D];
            const double c = E::template eval<B + 1, P, A0>(a, i, j, k);
            if (center) {
                const double l = E::template eval<B + 1, P, A0>(a, OPF_SHIFT(D, -1));
                return P::mul(P::div(P::sub(c, l), P::add(__ldg(ax.dx + q - 1), __ldg(ax.dx + q))), 2.0);
            } else {
                const double r = E::template eval<B + 1, P, A0>(a, OPF_SHIFT(D, 1));
                return P::div(P::sub(r, c), __ldg(ax.dx + q));
            }
        }
    };

    // D1FirstOrderBiasedDownwind<d>::eval (D1FirstOrderBiasedDownwind.hpp:53-57)
    template <int D, class E>
    struct D1Dn {
        static constexpr int size = 1 + E::size, maxaxis = (D > E::maxaxis ? D : E::maxaxis), nf = E::nf;
        template <int B, class P, bool A0>
        __device__ __forceinline__ static double eval(const ExprArgs& a, int i, int j, int k) {
            const bool center = (a.loc[B] >> D) & 1;
            const int q = axis_of<D>(i, j, k);
            const AxisView& ax = a.ax[D];
            const double c = E::template eval<B + 1, P, A0>(a, i, j, k);
            const double l = E::template eval<B + 1, P, A0>(a, OPF_SHIFT(D, -1));
            const double h = center ? P::mul(P::add(__ldg(ax.dx + q - 1), __ldg(ax.dx + q)), 0.5) : __ldg(ax.dx + q - 1);
            return P::div(P::sub(c, l), h);
        }
    };
    // D1FirstOrderBiasedUpwind<d>::eval (D1FirstOrderBiasedUpwind.hpp:54-58)
    template <int D, class E>
    struct D1Up {
        static constexpr int size = 1 + E::size, maxaxis = (D > E::maxaxis ? D : E::maxaxis), nf = E::nf;
        template <int B, class P, bool A0>
        __device__ __forceinline__ static double eval(const ExprArgs& a, int i, int j, int k) {
            const bool center = (a.loc[B] >> D) & 1;
            const int q = axis_of<D>(i, j, k);
            const AxisView& ax = a.ax[D];
            const double r = E::template eval<B + 1, P, A0>(a, OPF_SHIFT(D, 1));
            const double c = E::template eval<B + 1, P, A0>(a, i, j, k);
            const double h = center ? P::mul(P::add(__ldg(ax.dx + q), __ldg(ax.dx + q + 1)), 0.5) : __ldg(ax.dx + q);
            return P::div(P::sub(r, c), h);
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
            // a_k = g_k / e_k^2 (e_k can be ~1e-99 on flat data: keep the reference's form, no cross-products that
            // would underflow); one reciprocal of the sum instead of three divides
            const double a1 = .1 / (e1 * e1), a2 = .6 / (e2 * e2), a3 = .3 / (e3 * e3);
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

    // D1WENO53Downwind<d>::eval (D1WENO53Downwind.hpp:77-86): taps i-3..i+2, h = dx(d,i) ("uniform mesh is assumed")
    template <int D, class E>
    struct WenoDn {
        static constexpr int size = 1 + E::size, maxaxis = (D > E::maxaxis ? D : E::maxaxis), nf = E::nf;
        template <int B, class P, bool A0>
        __device__ __forceinline__ static double eval(const ExprArgs& a, int i, int j, int k) {
            const int q = axis_of<D>(i, j, k);
            const AxisView& ax = a.ax[D];
            const double pm3 = E::template eval<B + 1, P, A0>(a, OPF_SHIFT(D, -3));
            const double pm2 = E::template eval<B + 1, P, A0>(a, OPF_SHIFT(D, -2));
            const double pm1 = E::template eval<B + 1, P, A0>(a, OPF_SHIFT(D, -1));
            const double p0 = E::template eval<B + 1, P, A0>(a, i, j, k);
            const double pp1 = E::template eval<B + 1, P, A0>(a, OPF_SHIFT(D, 1));
            const double pp2 = E::template eval<B + 1, P, A0>(a, OPF_SHIFT(D, 2));
            if constexpr (P::fast) {
                const double rh = __ldg(ax.rdx + q);
                return weno53_core<P>((pm2 - pm3) * rh, (pm1 - pm2) * rh, (p0 - pm1) * rh, (pp1 - p0) * rh, (pp2 - pp1) * rh);
            } else {
                const double h = __ldg(ax.dx + q);
                return weno53_core<P>(P::div(P::sub(pm2, pm3), h), P::div(P::sub(pm1, pm2), h), P::div(P::sub(p0, pm1), h),
                                      P::div(P::sub(pp1, p0), h), P::div(P::sub(pp2, pp1), h));
            }
        }
    };
    // D1WENO53Upwind<d>::eval (D1WENO53Upwind.hpp:75-84): taps i-2..i+3
    template <int D, class E>
    struct WenoUp {
        static constexpr int size = 1 + E::size, maxaxis = (D > E::maxaxis ? D : E::maxaxis), nf = E::nf;
        template <int B, class P, bool A0>
        __device__ __forceinline__ static double eval(const ExprArgs& a, int i, int j, int k) {
            const int q = axis_of<D>(i, j, k);
            const AxisView& ax = a.ax[D];
            const double pm2 = E::template eval<B + 1, P, A0>(a, OPF_SHIFT(D, -2));
            const double pm1 = E::template eval<B + 1, P, A0>(a, OPF_SHIFT(D, -1));
            const double p0 = E::template eval<B + 1, P, A0>(a, i, j, k);
            const double pp1 = E::template eval<B + 1, P, A0>(a, OPF_SHIFT(D, 1));
            const double pp2 = E::template eval<B + 1, P, A0>(a, OPF_SHIFT(D, 2));
            const double pp3 = E::template eval<B + 1, P, A0>(a, OPF_SHIFT(D, 3));
            if constexpr (P::fast) {
                const double rh = __ldg(ax.rdx + q);
                return weno53_core<P>((pp3 - pp2) * rh, (pp2 - pp1) * rh, (pp1 - p0) * rh, (p0 - pm1) * rh, (pm1 - pm2) * rh);
            } else {
                const double h = __ldg(ax.dx + q);
                return weno53_core<P>(P::div(P::sub(pp3, pp2), h), P::div(P::sub(pp2, pp1), h), P::div(P::sub(pp1, p0), h),
                                      P::div(P::sub(p0, pm1), h), P::div(P::sub(pm1, pm2), h));
            }
        }
    };

    // D1Linear<d, Cen2Cor>::eval (D1Linear.hpp:36-42) with Interpolator1D::intp (Interpolator.hpp:21-23,38-40)
    template <int D, class E>
    struct IntpC2N {
        static constexpr int size = 1 + E::size, maxaxis = (D > E::maxaxis ? D : E::maxaxis), nf = E::nf;
        template <int B, class P, bool A0>
        __device__ __forceinline__ static double eval(const ExprArgs& a, int i, int j, int k) {
            const int q = axis_of<D>(i, j, k);
            const AxisView& ax = a.ax[D];
            const double x1 = P::add(__ldg(ax.x + q - 1), P::mul(0.5, __ldg(ax.dx + q - 1)));
            const double x2 = P::add(__ldg(ax.x + q), P::mul(0.5, __ldg(ax.dx + q)));
            const double y1 = E::template eval<B + 1, P, A0>(a, OPF_SHIFT(D, -1));
            const double y2 = E::template eval<B + 1, P, A0>(a, i, j, k);
            const double x = __ldg(ax.x + q);
            const double u1 = P::sub(x1, x), u2 = P::sub(x2, x);
            return P::div(P::sub(P::mul(u1, y2), P::mul(u2, y1)), P::sub(u1, u2));
        }
    };
    // D1Linear<d, Cor2Cen>::eval (D1Linear.hpp:44): Math::mid (Interpolator.hpp:83)
    template <int D, class E>
    struct IntpN2C {
        static constexpr int size = 1 + E::size, maxaxis = (D > E::maxaxis ? D : E::maxaxis), nf = E::nf;
        template <int B, class P, bool A0>
        __device__ __forceinline__ static double eval(const ExprArgs& a, int i, int j, int k) {
            const double y1 = E::template eval<B + 1, P, A0>(a, i, j, k);
            const double y2 = E::template eval<B + 1, P, A0>(a, OPF_SHIFT(D, 1));
            return P::mul(P::add(y1, y2), 0.5);
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
    template <class E, class P, bool A0, int DIM>
    __global__ void __launch_bounds__(512) assign_kernel(const __grid_constant__ ExprArgs a, const DstView dst,
                                                         const double* __restrict__ oldp, const LaunchRange r,
                                                         const int ch, const int op) {
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
            for (long long ii = b + threadIdx.x; ii < e; ii += blockDim.x)
                acc = red_comb(rop, acc, red_lift(rop, E::template eval<0, P, A0>(a, r.lo[0] + (int) ii, j, k)));
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
    inline LaunchGeom assign_geometry(const LaunchInfo& li) {
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
            const int ty = ety > 0 ? ety : 4;
            g.block = dim3(tx, ty, 1);
            g.ch = ech > 0 ? ech : 32;
            g.grid = dim3((n0 + tx - 1) / tx, (n1 + ty - 1) / ty, (n2 + g.ch - 1) / g.ch);
        }
        return g;
    }

    template <class E, class P, bool A0>
    int launch_assign(const ExprArgs& a, const LaunchInfo& li, cudaStream_t st) {
        const LaunchGeom g = assign_geometry(li);
        if (g.grid.x == 0) return 0;
        if constexpr (E::maxaxis < 1)
            if (li.dim == 1) {
                assign_kernel<E, P, A0, 1><<<g.grid, g.block, 0, st>>>(a, li.dst, li.old, li.r, g.ch, li.op);
                return (int) cudaGetLastError();
            }
        if constexpr (E::maxaxis < 2)
            if (li.dim == 2) {
                assign_kernel<E, P, A0, 2><<<g.grid, g.block, 0, st>>>(a, li.dst, li.old, li.r, g.ch, li.op);
                return (int) cudaGetLastError();
            }
        if (li.dim == 3) {
            assign_kernel<E, P, A0, 3><<<g.grid, g.block, 0, st>>>(a, li.dst, li.old, li.r, g.ch, li.op);
            return (int) cudaGetLastError();
        }
        return -2;// expression uses an axis the field does not have
    }
    template <class E, class P, bool A0>
    int launch_reduce(const ExprArgs& a, const LaunchInfo& li, cudaStream_t st) {
        const int nb = li.n_partials;
        reduce_kernel<E, P, A0><<<nb, 256, 0, st>>>(a, li.r, li.partials, li.rop);
        reduce_final_kernel<0><<<1, 256, 0, st>>>(li.partials, nb, li.partials + nb, li.rop);
        return (int) cudaGetLastError();
    }

    // LaunchInfo.rop < 0 -> assignment; >= 0 -> reduction
    template <class E>
    int launcher(const void* args_blob, const void* launch_blob, void* stream) {
        const ExprArgs& a = *static_cast<const ExprArgs*>(args_blob);
        const LaunchInfo& li = *static_cast<const LaunchInfo*>(launch_blob);
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        constexpr bool can_alias = E::nf > 1;
        if (li.rop >= 0) {
            if constexpr (can_alias)
                if (li.alias0) return li.mode == 0 ? launch_reduce<E, Exact, true>(a, li, st) : launch_reduce<E, Fast, true>(a, li, st);
            return li.mode == 0 ? launch_reduce<E, Exact, false>(a, li, st) : launch_reduce<E, Fast, false>(a, li, st);
        }
        if constexpr (can_alias)
            if (li.alias0) return li.mode == 0 ? launch_assign<E, Exact, true>(a, li, st) : launch_assign<E, Fast, true>(a, li, st);
        return li.mode == 0 ? launch_assign<E, Exact, false>(a, li, st) : launch_assign<E, Fast, false>(a, li, st);
    }

}// namespace opf
