// opflow/field.hpp -- expression templates, operators, CartesianField and ExprBuilder of the B200 front-end.
//
// Reference spellings kept:
//   Expression<Op, Args...>            src/Core/Expr/Expression.hpp:24-111, makeExpression :113-127
//   scalar wrapping (ScalarExpr)       src/Core/Expr/ScalarExpr.hpp:23-60, BinOpDefMacros.hpp.in:159-191
//   d1/dx/dy/dz, d2/d2x/d2y/d2z        src/Core/Operator/FDMOperators/DiffsInterface.hpp:21-47
//   d1IntpCenterToCorner / CornerToCenter   src/Core/Operator/Interpolator/IntpInterface.hpp:26-40
//   conditional                        src/Core/Operator/Conditional.hpp:72-76
//   CartesianField / ExprBuilder       src/Core/Field/MeshBased/Structured/CartesianField.hpp:37-1041
//   FieldAssigner::assign              src/Core/Loops/FieldAssigner.hpp:26-86
//
// What differs underneath: an expression is never walked cell by cell on the host.  Its *type* is mapped to a device functor
// type (opf::Add<opf::F<0>, ...>, opflow_b200/csrc/opf_device.cuh); assignment flattens the leaves (field handles, scalars) in
// preorder and makes one opf_assign call.  Compiled with nvcc the kernels of every assigned expression type are instantiated in
// the user's translation unit and registered by signature (opf_expr_register).
#pragma once
#include "devtypes.hpp"
#include "mesh.hpp"

namespace OpFlow {
    struct ExprTag {};
    template <typename T>
    concept ExprType = std::is_base_of_v<ExprTag, std::remove_cvref_t<T>>;

    template <typename D, typename M>
    struct CartesianField;
    template <typename F>
    struct ExprBuilder;

    namespace internal {
        template <typename T>
        struct is_cartesian_field : std::false_type {};
        template <typename D, typename M>
        struct is_cartesian_field<CartesianField<D, M>> : std::true_type {};
        template <typename T>
        concept FieldType = is_cartesian_field<std::remove_cvref_t<T>>::value;

        // leaves of a flattened expression, in preorder (== the numbering of opf::F<k> / opf::S<k> in the device type)
        struct Flat {
            std::vector<opf_field_t> fields;
            std::vector<double> scalars;
            unsigned mask = 0;// bit k: field leaf k is the equation's unknown
            std::string sig;
        };
    }// namespace internal

    // ScalarExpr<T>::evalAt ignores the index (ScalarExpr.hpp:36)
    struct ScalarExpr : ExprTag {
        Real val = 0;
        static constexpr bool has_unknown = false;
        static constexpr int dim = 0;
        template <int NF, int NS>
        struct Dev {
            using type = opf::S<NS>;
            static constexpr int nf = NF, ns = NS + 1;
        };
        void flatten(internal::Flat& fl) const {
            fl.sig += "S<" + std::to_string(fl.scalars.size()) + ">";
            fl.scalars.push_back(val);
        }
    };

    // ExprProxy of an lvalue field (Expr.hpp:322-336): the tree refers to the field, it never copies it
    template <typename F>
    struct FieldRef : ExprTag {
        const F* f = nullptr;
        static constexpr bool has_unknown = false;
        static constexpr int dim = F::dim;
        template <int NF, int NS>
        struct Dev {
            using type = opf::F<NF>;
            static constexpr int nf = NF + 1, ns = NS;
        };
        void flatten(internal::Flat& fl) const {
            f->syncToDevice();
            fl.sig += "F<" + std::to_string(fl.fields.size()) + ">";
            fl.fields.push_back(f->h);
        }
    };

    // the `e` handed to an equation lambda (reference: StencilField over the target, StencilField.hpp:30-120): here simply a
    // field leaf flagged as the unknown -- the engine applies the expression to its Krylov vectors instead of assembling rows
    template <typename F>
    struct UnknownRef : ExprTag {
        const F* target = nullptr;
        static constexpr bool has_unknown = true;
        static constexpr int dim = F::dim;
        template <int NF, int NS>
        struct Dev {
            using type = opf::F<NF>;
            static constexpr int nf = NF + 1, ns = NS;
        };
        void flatten(internal::Flat& fl) const {
            fl.mask |= 1u << fl.fields.size();
            fl.sig += "F<" + std::to_string(fl.fields.size()) + ">";
            fl.fields.push_back(target->h);
        }
    };

    namespace internal {
        // device types of a pack of children, numbered left to right (preorder)
        template <int NF, int NS, typename... A>
        struct DevPack;
        template <int NF, int NS>
        struct DevPack<NF, NS> {
            static constexpr int nf = NF, ns = NS;
            template <template <class...> class Node>
            using apply = void;
        };
        template <int NF, int NS, typename A0>
        struct DevPack<NF, NS, A0> {
            using D0 = typename A0::template Dev<NF, NS>;
            static constexpr int nf = D0::nf, ns = D0::ns;
        };
        template <int NF, int NS, typename A0, typename A1>
        struct DevPack<NF, NS, A0, A1> {
            using D0 = typename A0::template Dev<NF, NS>;
            using D1 = typename A1::template Dev<D0::nf, D0::ns>;
            static constexpr int nf = D1::nf, ns = D1::ns;
        };
        template <int NF, int NS, typename A0, typename A1, typename A2>
        struct DevPack<NF, NS, A0, A1, A2> {
            using D0 = typename A0::template Dev<NF, NS>;
            using D1 = typename A1::template Dev<D0::nf, D0::ns>;
            using D2 = typename A2::template Dev<D1::nf, D1::ns>;
            static constexpr int nf = D2::nf, ns = D2::ns;
        };
    }// namespace internal

    template <typename Op, typename... Args>
    struct Expression : ExprTag {
        std::tuple<Args...> args;
        static constexpr bool has_unknown = (Args::has_unknown || ...);
        static constexpr int dim = std::max({Args::dim...});
        explicit Expression(Args... a) : args(std::move(a)...) {}
        // Expr::prepare() (Expression.hpp:99-103 + every Op::prepare): ranges and location of the result, computed by the
        // engine's bit-exact range algebra (opf_expr_prepare) -- host only, no device work
        mutable DS::Range<(dim > 0 ? dim : 1)> accessibleRange, localRange, logicalRange, assignableRange;
        mutable std::array<LocOnMesh, (dim > 0 ? dim : 1)> loc {};
        void prepare() const {
            if constexpr (dim > 0) {
                internal::Flat fl;
                flatten(fl);
                opf_range r;
                int l[OPF_MAX_DIM];
                internal::check_rc(opf_expr_prepare(fl.sig.c_str(), fl.fields.data(), (int) fl.fields.size(), 2, &r, l), "opf_expr_prepare");
                accessibleRange = internal::from_c<dim>(r);
                internal::check_rc(opf_expr_prepare(fl.sig.c_str(), fl.fields.data(), (int) fl.fields.size(), 0, &r, nullptr), "opf_expr_prepare");
                localRange = internal::from_c<dim>(r);
                internal::check_rc(opf_expr_prepare(fl.sig.c_str(), fl.fields.data(), (int) fl.fields.size(), 3, &r, nullptr), "opf_expr_prepare");
                logicalRange = internal::from_c<dim>(r);
                assignableRange.setEmpty();
                for (int d = 0; d < dim; ++d) loc[d] = static_cast<LocOnMesh>(l[d]);
            }
        }
        template <int NF, int NS>
        struct Dev {
            using P = internal::DevPack<NF, NS, Args...>;
            static constexpr int nf = P::nf, ns = P::ns;
            static constexpr auto pick() {
                if constexpr (sizeof...(Args) == 1) return std::type_identity<typename Op::template dev<typename P::D0::type>> {};
                else if constexpr (sizeof...(Args) == 2)
                    return std::type_identity<typename Op::template dev<typename P::D0::type, typename P::D1::type>> {};
                else
                    return std::type_identity<typename Op::template dev<typename P::D0::type, typename P::D1::type, typename P::D2::type>> {};
            }
            using type = typename decltype(pick())::type;
        };
        void flatten(internal::Flat& fl) const {
            fl.sig += Op::name;
            // adaptor ops carry "Adapt1<functorName" as their name: the operands follow after a comma
            const bool adaptor = fl.sig.size() >= 1 && std::string_view(Op::name).find('<') != std::string_view::npos;
            fl.sig += adaptor ? "," : "<";
            if constexpr (Op::axis >= 0) fl.sig += std::to_string(Op::axis) + ",";
            bool first = true;
            std::apply(
                    [&](const auto&... a) {
                        ((fl.sig += (first ? "" : ","), first = false, a.flatten(fl)), ...);
                    },
                    args);
            fl.sig += ">";
        }
    };

    namespace internal {
        template <typename T>
        auto wrap(T&& t) {
            using R = std::remove_cvref_t<T>;
            if constexpr (is_cartesian_field<R>::value) return FieldRef<R> {{}, &t};
            else if constexpr (std::is_arithmetic_v<R>)
                return ScalarExpr {{}, static_cast<Real>(t)};
            else
                return R(std::forward<T>(t));
        }
        template <typename T>
        using wrapped_t = decltype(wrap(std::declval<T>()));

        template <typename A, typename B>
        concept ExprOperands = (ExprType<A> || ExprType<B>) && (ExprType<A> || Meta::Numerical<A>) && (ExprType<B> || Meta::Numerical<B>);
    }// namespace internal

    template <typename Op, typename... T>
    auto makeExpression(T&&... t) {
        return Expression<Op, internal::wrapped_t<T>...>(internal::wrap(std::forward<T>(t))...);
    }

    // ------------------------------------------------------------------------------------------------ point-wise operators
    // AMDS.hpp / Compare.hpp / Boolean.hpp / MinMax.hpp / Conditional.hpp: `name` is the node name of the signature grammar
#define OPF_FE_OP2(OpName, Node)                                                                                       \
    struct OpName {                                                                                                    \
        static constexpr const char* name = #Node;                                                                     \
        static constexpr int axis = -1;                                                                                \
        template <class A, class B>                                                                                    \
        using dev = opf::Node<A, B>;                                                                                   \
    };
#define OPF_FE_OP1(OpName, Node)                                                                                       \
    struct OpName {                                                                                                    \
        static constexpr const char* name = #Node;                                                                     \
        static constexpr int axis = -1;                                                                                \
        template <class A>                                                                                             \
        using dev = opf::Node<A>;                                                                                      \
    };
    OPF_FE_OP2(AddOp, Add)
    OPF_FE_OP2(SubOp, Sub)
    OPF_FE_OP2(MulOp, Mul)
    OPF_FE_OP2(DivOp, Div)
    OPF_FE_OP2(MinOp, Min)
    OPF_FE_OP2(MaxOp, Max)
    OPF_FE_OP2(PowOp, Pow)
    OPF_FE_OP2(LessThanOp, Lt)
    OPF_FE_OP2(LessThanOrEqualOp, Le)
    OPF_FE_OP2(GreaterThanOp, Gt)
    OPF_FE_OP2(GreaterThanOrEqualOp, Ge)
    OPF_FE_OP2(NotEqualToOp, Ne)
    OPF_FE_OP2(AndOp, And)
    OPF_FE_OP2(OrOp, Or)
    OPF_FE_OP1(NegOp, Neg)
    OPF_FE_OP1(PosOp, Pos)
    OPF_FE_OP1(NotOp, Not)
    OPF_FE_OP1(SqrtOp, Sqrt)
    OPF_FE_OP1(AbsOp, Abs)
    OPF_FE_OP1(ExpOp, Exp)
    OPF_FE_OP1(LogOp, Log)
    OPF_FE_OP1(SinOp, Sin)
    OPF_FE_OP1(CosOp, Cos)
    OPF_FE_OP1(TanOp, Tan)
    OPF_FE_OP1(TanhOp, Tanh)
    OPF_FE_OP1(Pow2Op, Pow2)
    OPF_FE_OP1(Exp2Op, Exp2)
    OPF_FE_OP1(Expm1Op, Expm1)
    OPF_FE_OP1(Log10Op, Log10)
    OPF_FE_OP1(Log2Op, Log2)
    OPF_FE_OP1(Log1pOp, Log1p)
    OPF_FE_OP1(CbrtOp, Cbrt)
    OPF_FE_OP1(ASinOp, ASin)
    OPF_FE_OP1(ACosOp, ACos)
    OPF_FE_OP1(ATanOp, ATan)
    OPF_FE_OP1(SinhOp, Sinh)
    OPF_FE_OP1(CoshOp, Cosh)
    OPF_FE_OP1(ASinhOp, ASinh)
    OPF_FE_OP1(ACoshOp, ACosh)
    OPF_FE_OP1(ATanhOp, ATanh)
    OPF_FE_OP1(ErfOp, Erf)
    OPF_FE_OP1(ErfcOp, Erfc)
    OPF_FE_OP1(TGammaOp, TGamma)
    OPF_FE_OP1(LGammaOp, LGamma)
    OPF_FE_OP1(CeilOp, Ceil)
    OPF_FE_OP1(FloorOp, Floor)
    OPF_FE_OP1(TruncOp, Trunc)
    OPF_FE_OP1(RoundOp, Round)
    OPF_FE_OP1(LRoundOp, LRound)
    OPF_FE_OP1(LLRoundOp, LLRound)
    OPF_FE_OP1(NearbyIntOp, NearbyInt)
    OPF_FE_OP1(RintOp, Rint)
    OPF_FE_OP1(LRintOp, LRint)
    OPF_FE_OP1(LLRintOp, LLRint)
    OPF_FE_OP1(ILogbOp, ILogb)
    OPF_FE_OP1(LogbOp, Logb)
    OPF_FE_OP2(FModOp, FMod)
    OPF_FE_OP2(RemainderOp, Remainder)
    OPF_FE_OP2(FDimOp, FDim)
    OPF_FE_OP2(HypotOp, Hypot)
    OPF_FE_OP2(ATan2Op, ATan2)
    OPF_FE_OP2(LdexpOp, Ldexp)
    OPF_FE_OP2(ScalbnOp, Scalbn)
    OPF_FE_OP2(ScalblnOp, Scalbln)
    OPF_FE_OP2(NextafterOp, Nextafter)
    OPF_FE_OP2(NexttowardOp, Nexttoward)
    OPF_FE_OP2(CopysingOp, Copysing)
#undef OPF_FE_OP1
#undef OPF_FE_OP2
    struct CondOp {
        static constexpr const char* name = "Cond";
        static constexpr int axis = -1;
        template <class C, class A, class B>
        using dev = opf::Cond<C, A, B>;
    };

#define OPF_FE_BINARY(sym, OpName)                                                                                     \
    template <typename A, typename B>                                                                                  \
    requires internal::ExprOperands<A, B> auto operator sym(A&& a, B&& b) {                                            \
        return makeExpression<OpName>(std::forward<A>(a), std::forward<B>(b));                                         \
    }
    OPF_FE_BINARY(+, AddOp)
    OPF_FE_BINARY(-, SubOp)
    OPF_FE_BINARY(*, MulOp)
    OPF_FE_BINARY(/, DivOp)
    OPF_FE_BINARY(<, LessThanOp)
    OPF_FE_BINARY(<=, LessThanOrEqualOp)
    OPF_FE_BINARY(>, GreaterThanOp)
    OPF_FE_BINARY(>=, GreaterThanOrEqualOp)
    OPF_FE_BINARY(!=, NotEqualToOp)
    OPF_FE_BINARY(&&, AndOp)
    OPF_FE_BINARY(||, OrOp)
#undef OPF_FE_BINARY
    template <ExprType A>
    auto operator-(A&& a) {
        return makeExpression<NegOp>(std::forward<A>(a));
    }
    template <ExprType A>
    auto operator+(A&& a) {
        return makeExpression<PosOp>(std::forward<A>(a));
    }
    template <ExprType A>
    auto operator!(A&& a) {
        return makeExpression<NotOp>(std::forward<A>(a));
    }
#define OPF_FE_FUNC1(fname, OpName)                                                                                    \
    template <ExprType A>                                                                                              \
    auto fname(A&& a) {                                                                                                \
        return makeExpression<OpName>(std::forward<A>(a));                                                             \
    }
    OPF_FE_FUNC1(sqrt, SqrtOp)
    OPF_FE_FUNC1(abs, AbsOp)
    OPF_FE_FUNC1(exp, ExpOp)
    OPF_FE_FUNC1(log, LogOp)
    OPF_FE_FUNC1(sin, SinOp)
    OPF_FE_FUNC1(cos, CosOp)
    OPF_FE_FUNC1(tan, TanOp)
    OPF_FE_FUNC1(tanh, TanhOp)
    OPF_FE_FUNC1(pow2, Pow2Op)
    OPF_FE_FUNC1(exp2, Exp2Op)
    OPF_FE_FUNC1(expm1, Expm1Op)
    OPF_FE_FUNC1(log10, Log10Op)
    OPF_FE_FUNC1(log2, Log2Op)
    OPF_FE_FUNC1(log1p, Log1pOp)
    OPF_FE_FUNC1(cbrt, CbrtOp)
    OPF_FE_FUNC1(asin, ASinOp)
    OPF_FE_FUNC1(acos, ACosOp)
    OPF_FE_FUNC1(atan, ATanOp)
    OPF_FE_FUNC1(sinh, SinhOp)
    OPF_FE_FUNC1(cosh, CoshOp)
    OPF_FE_FUNC1(asinh, ASinhOp)
    OPF_FE_FUNC1(acosh, ACoshOp)
    OPF_FE_FUNC1(atanh, ATanhOp)
    OPF_FE_FUNC1(erf, ErfOp)
    OPF_FE_FUNC1(erfc, ErfcOp)
    OPF_FE_FUNC1(tgamma, TGammaOp)
    OPF_FE_FUNC1(lgamma, LGammaOp)
    OPF_FE_FUNC1(ceil, CeilOp)
    OPF_FE_FUNC1(floor, FloorOp)
    OPF_FE_FUNC1(trunc, TruncOp)
    OPF_FE_FUNC1(round, RoundOp)
    OPF_FE_FUNC1(lround, LRoundOp)
    OPF_FE_FUNC1(llround, LLRoundOp)
    OPF_FE_FUNC1(nearbyint, NearbyIntOp)
    OPF_FE_FUNC1(rint, RintOp)
    OPF_FE_FUNC1(lrint, LRintOp)
    OPF_FE_FUNC1(llrint, LLRintOp)
    OPF_FE_FUNC1(ilogb, ILogbOp)
    OPF_FE_FUNC1(logb, LogbOp)
#undef OPF_FE_FUNC1
#define OPF_FE_FUNC2(fname, OpName)                                                                                    \
    template <typename A, typename B>                                                                                  \
    requires internal::ExprOperands<A, B> auto fname(A&& a, B&& b) {                                                   \
        return makeExpression<OpName>(std::forward<A>(a), std::forward<B>(b));                                         \
    }
    OPF_FE_FUNC2(fmod, FModOp)
    OPF_FE_FUNC2(remainder, RemainderOp)
    OPF_FE_FUNC2(fdim, FDimOp)
    OPF_FE_FUNC2(hypot, HypotOp)
    OPF_FE_FUNC2(atan2, ATan2Op)
    OPF_FE_FUNC2(ldexp, LdexpOp)
    OPF_FE_FUNC2(scalbn, ScalbnOp)
    OPF_FE_FUNC2(scalbln, ScalblnOp)
    OPF_FE_FUNC2(nextafter, NextafterOp)
    OPF_FE_FUNC2(nexttoward, NexttowardOp)
    OPF_FE_FUNC2(copysign, CopysingOp)
#undef OPF_FE_FUNC2
    template <typename A, typename B>
    requires internal::ExprOperands<A, B> auto min(A&& a, B&& b) {
        return makeExpression<MinOp>(std::forward<A>(a), std::forward<B>(b));
    }
    template <typename A, typename B>
    requires internal::ExprOperands<A, B> auto max(A&& a, B&& b) {
        return makeExpression<MaxOp>(std::forward<A>(a), std::forward<B>(b));
    }
    template <typename A, typename B>
    requires internal::ExprOperands<A, B> auto pow(A&& a, B&& b) {
        return makeExpression<PowOp>(std::forward<A>(a), std::forward<B>(b));
    }
    template <typename C, typename A, typename B>
    requires(ExprType<C> || ExprType<A> || ExprType<B>) auto conditional(C&& c, A&& a, B&& b) {
        return makeExpression<CondOp>(std::forward<C>(c), std::forward<A>(a), std::forward<B>(b));
    }

    // ------------------------------------------------------------------------------------------------ stencil operators
#define OPF_FE_STENCIL(KernelName, Node, width)                                                                        \
    template <std::size_t d>                                                                                           \
    struct KernelName {                                                                                                \
        static constexpr const char* name = #Node;                                                                     \
        static constexpr int axis = static_cast<int>(d);                                                               \
        static constexpr int bc_width = width;                                                                         \
        template <class A>                                                                                             \
        using dev = opf::Node<static_cast<int>(d), A>;                                                                 \
    };
    OPF_FE_STENCIL(D2SecondOrderCentered, D2C, 1)        // D2SecondOrderCentered.hpp:22-272
    OPF_FE_STENCIL(D1FirstOrderCentered, D1C, 1)         // D1FirstOrderCentered.hpp:20-110
    OPF_FE_STENCIL(D1FirstOrderBiasedDownwind, D1Dn, 1)  // D1FirstOrderBiasedDownwind.hpp:21-170
    OPF_FE_STENCIL(D1FirstOrderBiasedUpwind, D1Up, 1)    // D1FirstOrderBiasedUpwind.hpp:22-170
    OPF_FE_STENCIL(D1WENO53Downwind, WenoDn, 3)          // D1WENO53Downwind.hpp:24-200
    OPF_FE_STENCIL(D1WENO53Upwind, WenoUp, 3)            // D1WENO53Upwind.hpp:24-200
#undef OPF_FE_STENCIL
    enum class IntpDirection { Cor2Cen, Cen2Cor };
    template <std::size_t d, IntpDirection dir>
    struct D1Linear {// D1Linear.hpp:24-90
        static constexpr const char* name = dir == IntpDirection::Cen2Cor ? "IntpC2N" : "IntpN2C";
        static constexpr int axis = static_cast<int>(d);
        static constexpr int bc_width = 1;
        template <class A>
        using dev = std::conditional_t<dir == IntpDirection::Cen2Cor, opf::IntpC2N<static_cast<int>(d), A>, opf::IntpN2C<static_cast<int>(d), A>>;
    };

    template <typename Kernel, typename E>
    auto d1(E&& expr) {
        return makeExpression<Kernel>(std::forward<E>(expr));
    }
    template <typename Kernel, typename E>
    auto d2(E&& expr) {
        return makeExpression<Kernel>(std::forward<E>(expr));
    }
#define OPF_FE_THREED(x_name, y_name, z_name, u_name)                                                                  \
    template <template <std::size_t> typename Kernel, typename E>                                                      \
    auto x_name(E&& expr) {                                                                                            \
        return u_name<Kernel<0>>(std::forward<E>(expr));                                                               \
    }                                                                                                                  \
    template <template <std::size_t> typename Kernel, typename E>                                                      \
    auto y_name(E&& expr) {                                                                                            \
        return u_name<Kernel<1>>(std::forward<E>(expr));                                                               \
    }                                                                                                                  \
    template <template <std::size_t> typename Kernel, typename E>                                                      \
    auto z_name(E&& expr) {                                                                                            \
        return u_name<Kernel<2>>(std::forward<E>(expr));                                                               \
    }
    OPF_FE_THREED(dx, dy, dz, d1)
    OPF_FE_THREED(d2x, d2y, d2z, d2)
#undef OPF_FE_THREED
    // IntpInterface.hpp:28-44: the operands are forwarded as they come -- (e) for D1Linear, (u, e) for the flux-limiter kernels
    template <std::size_t dim, template <std::size_t, IntpDirection> typename Kernel = D1Linear, typename... E>
    auto d1IntpCenterToCorner(E&&... expr) {
        return makeExpression<Kernel<dim, IntpDirection::Cen2Cor>>(std::forward<E>(expr)...);
    }
    template <std::size_t dim, template <std::size_t, IntpDirection> typename Kernel = D1Linear, typename... E>
    auto d1IntpCornerToCenter(E&&... expr) {
        return makeExpression<Kernel<dim, IntpDirection::Cor2Cen>>(std::forward<E>(expr)...);
    }
    template <std::size_t dim, IntpDirection dir, template <std::size_t, IntpDirection> typename Kernel = D1Linear, typename... E>
    auto d1Intp(E&&... expr) {
        return makeExpression<Kernel<dim, dir>>(std::forward<E>(expr)...);
    }

    // ------------------------------------------------------------------------------------------------ flux-limiter interpolators
    // FluxLimiterKernels.hpp:22-80 + D1FluxLimiter.hpp + D1FluxLimiterBasedIntpOp.hpp:22-61.  The kernel TYPES only select the device
    // node (opf::Fl<Scheme>C2N / N2C, opf_device.cuh); the generalised piecewise-linear families (SPL / GPL) have no device node yet.
    enum class KappaScheme { CDS, QUICK, CUI, Fromm, LUI };
    template <KappaScheme s>
    struct KappaKernel {};
    struct MinmodKernel {};
    struct SuperbeeKernel {};
    struct MUSCLKernel {};
    struct HarmonicKernel {};
    struct vanAlbadaKernel {};
    namespace internal {
        template <typename K>
        struct FlNode;
#define OPF_FE_FLNODE(KernelType, Node)                                                                                \
    template <>                                                                                                        \
    struct FlNode<KernelType> {                                                                                        \
        static constexpr const char *c2n = #Node "C2N", *n2c = #Node "N2C";                                            \
        template <int D, class U, class E>                                                                             \
        using C2N = opf::Node##C2N<D, U, E>;                                                                           \
        template <int D, class U, class E>                                                                             \
        using N2C = opf::Node##N2C<D, U, E>;                                                                           \
    };
        OPF_FE_FLNODE(KappaKernel<KappaScheme::CDS>, FlCentral)
        OPF_FE_FLNODE(KappaKernel<KappaScheme::QUICK>, FlQuick)
        OPF_FE_FLNODE(KappaKernel<KappaScheme::CUI>, FlCui)
        OPF_FE_FLNODE(KappaKernel<KappaScheme::Fromm>, FlFromm)
        OPF_FE_FLNODE(KappaKernel<KappaScheme::LUI>, FlLui)
        OPF_FE_FLNODE(MinmodKernel, FlMinmod)
        OPF_FE_FLNODE(SuperbeeKernel, FlSuperbee)
        OPF_FE_FLNODE(MUSCLKernel, FlMuscl)
        OPF_FE_FLNODE(HarmonicKernel, FlHarmonic)
        OPF_FE_FLNODE(vanAlbadaKernel, FlAlbada)
#undef OPF_FE_FLNODE
        template <typename K, std::size_t d, IntpDirection dir>
        struct D1FluxLimiterImpl {// D1FluxLimiter.hpp:137-203: operands (u, e)
            static constexpr const char* name = dir == IntpDirection::Cen2Cor ? FlNode<K>::c2n : FlNode<K>::n2c;
            static constexpr int axis = static_cast<int>(d);
            static constexpr int bc_width = 2;
            template <class U, class E>
            using dev = std::conditional_t<dir == IntpDirection::Cen2Cor, typename FlNode<K>::template C2N<static_cast<int>(d), U, E>,
                                           typename FlNode<K>::template N2C<static_cast<int>(d), U, E>>;
        };
    }// namespace internal
    template <typename Kernel>
    struct D1FluxLimiterGen {
        template <std::size_t d, IntpDirection dir>
        using Op = internal::D1FluxLimiterImpl<Kernel, d, dir>;
    };
    template <std::size_t d, IntpDirection dir>
    using D1QUICK = typename D1FluxLimiterGen<KappaKernel<KappaScheme::QUICK>>::template Op<d, dir>;
    template <std::size_t d, IntpDirection dir>
    using D1Central = typename D1FluxLimiterGen<KappaKernel<KappaScheme::CDS>>::template Op<d, dir>;
    template <std::size_t d, IntpDirection dir>
    using D1CUI = typename D1FluxLimiterGen<KappaKernel<KappaScheme::CUI>>::template Op<d, dir>;
    template <std::size_t d, IntpDirection dir>
    using D1Fromm = typename D1FluxLimiterGen<KappaKernel<KappaScheme::Fromm>>::template Op<d, dir>;
    template <std::size_t d, IntpDirection dir>
    using D1LinearUpwind = typename D1FluxLimiterGen<KappaKernel<KappaScheme::LUI>>::template Op<d, dir>;
    template <std::size_t d, IntpDirection dir>
    using D1Minmod = typename D1FluxLimiterGen<MinmodKernel>::template Op<d, dir>;
    template <std::size_t d, IntpDirection dir>
    using D1Superbee = typename D1FluxLimiterGen<SuperbeeKernel>::template Op<d, dir>;
    template <std::size_t d, IntpDirection dir>
    using D1MUSCL = typename D1FluxLimiterGen<MUSCLKernel>::template Op<d, dir>;
    template <std::size_t d, IntpDirection dir>
    using D1Harmonic = typename D1FluxLimiterGen<HarmonicKernel>::template Op<d, dir>;
    template <std::size_t d, IntpDirection dir>
    using D1Albada = typename D1FluxLimiterGen<vanAlbadaKernel>::template Op<d, dir>;

    // ------------------------------------------------------------------------------------------------ element-wise functor adaptors
    // Utils::NamedFunctor / makeCXprString (src/Utils/NamedFunctor.hpp:21-52, ConstexprString.hpp) and UniOpAdaptor / BinOpAdaptor
    // (src/Core/Operator/PerElemOpAdaptor.hpp:23-110).  The functor runs inside the device kernels: its call operator must be usable in
    // device code (a `__host__ __device__` function object, or a constexpr one under --expt-relaxed-constexpr).  Its name becomes part
    // of the expression signature, so two functors never share a kernel.  nvcc cannot name a lambda's closure type in a kernel's
    // template arguments, so the functor has to be an object of a NAMED type (struct SmoothDelta { OPF_HD double operator()(double) const; });
    // the `constexpr auto func = [=](Real d) {...}` form of UniLS.cpp:109 compiles against the reference only.
    namespace Utils {
        template <std::size_t N>
        struct CXprString {
            char s[N] {};
            constexpr CXprString(const char (&str)[N]) {
                for (std::size_t i = 0; i < N; ++i) s[i] = str[i];
            }
            constexpr std::string to_string() const { return std::string(s); }
        };
        template <std::size_t N>
        constexpr auto makeCXprString(const char (&str)[N]) {
            return CXprString<N>(str);
        }
        template <auto Functor, auto Name>
        struct NamedFunctor {
            constexpr NamedFunctor() = default;
            constexpr static auto getFunc() { return Functor; }
            constexpr static auto getName() { return Name; }
            template <typename... A>
#if defined(__CUDACC__)
            __host__ __device__
#endif
                    constexpr auto
                    operator()(A&&... args) const {
                return Functor(std::forward<A>(args)...);
            }
        };
        template <typename T>
        struct is_named_functor : std::false_type {};
        template <auto F, auto N>
        struct is_named_functor<NamedFunctor<F, N>> : std::true_type {};
        template <typename T>
        concept NamedFunctorType = is_named_functor<std::remove_cvref_t<T>>::value;
    }// namespace Utils
    namespace internal {
        template <auto Functor>
        inline const std::string adaptor_name = [] {
            std::string n;
            for (char c : Functor.getName().to_string())
                if (c) n.push_back(std::isalnum((unsigned char) c) ? c : '_');
            return n.empty() ? std::string("unnamed_functor") : n;
        }();
    }
    template <Utils::NamedFunctorType auto Functor>
    struct UniOpAdaptor {
        static inline const std::string name_str = "Adapt1<" + internal::adaptor_name<Functor>;
        static inline const char* name = name_str.c_str();// flatten() appends "<" after the op name: the functor name follows as "Adapt1<name,...": see below
        static constexpr int axis = -1;
        static constexpr int bc_width = 0;
        template <class A>
        using dev = opf::Adapt1<std::remove_cvref_t<decltype(Functor.getFunc())>, A>;
    };
    template <Utils::NamedFunctorType auto Functor>
    struct BinOpAdaptor {
        static inline const std::string name_str = "Adapt2<" + internal::adaptor_name<Functor>;
        static inline const char* name = name_str.c_str();
        static constexpr int axis = -1;
        static constexpr int bc_width = 0;
        template <class A, class B>
        using dev = opf::Adapt2<std::remove_cvref_t<decltype(Functor.getFunc())>, A, B>;
    };

    // The nvcc-proof spelling: the functor is a named TYPE carrying its own name --
    //   struct SmoothDelta { static constexpr const char* name = "smoothDelta"; OPF_HD double operator()(double d) const {...} };
    //   auto delta = makeExpression<UniOpAdaptorT<SmoothDelta>>(p0);
    // (nvcc 12.9 can carry neither a lambda's closure type nor a class-type non-type template argument -- NamedFunctor<func, name> is
    // both -- through the host stubs of a kernel template, so UniOpAdaptor<functor> above only compiles in host-only builds.)
    namespace internal {
        template <typename Fn>
        inline const std::string adaptor_type_name = [] {
            std::string n;
            for (const char* c = Fn::name; *c; ++c) n.push_back(std::isalnum((unsigned char) *c) ? *c : '_');
            return n;
        }();
    }
    template <typename Fn>
    struct UniOpAdaptorT {
        static inline const std::string name_str = "Adapt1<" + internal::adaptor_type_name<Fn>;
        static inline const char* name = name_str.c_str();
        static constexpr int axis = -1;
        static constexpr int bc_width = 0;
        template <class A>
        using dev = opf::Adapt1<Fn, A>;
    };
    template <typename Fn>
    struct BinOpAdaptorT {
        static inline const std::string name_str = "Adapt2<" + internal::adaptor_type_name<Fn>;
        static inline const char* name = name_str.c_str();
        static constexpr int axis = -1;
        static constexpr int bc_width = 0;
        template <class A, class B>
        using dev = opf::Adapt2<Fn, A, B>;
    };

    // ------------------------------------------------------------------------------------------------ convolution
    // conv(expr, kernel) (Convolution.hpp:160-170): the kernel tensor's entries travel as consecutive scalar leaves (x fastest)
    namespace DS {
        template <typename T, int... ns>
        struct FixedSizeTensor {// DataStructures/Arrays/Tensor/FixedSizeTensor.hpp: dense, first index fastest
            static constexpr int dim = sizeof...(ns);
            static constexpr std::array<int, sizeof...(ns)> dims {ns...};
            static constexpr int total = (ns * ...);
            std::array<T, total> val {};
            constexpr FixedSizeTensor() = default;
            template <typename... V>
            requires(sizeof...(V) >= 1 && (std::is_convertible_v<V, T> && ...)) constexpr FixedSizeTensor(V... v) {
                if constexpr (sizeof...(V) == 1) val.fill((static_cast<T>(v), ...));
                else
                    val = {static_cast<T>(v)...};
            }
            constexpr int offset(const auto& idx) const {
                int o = 0, s = 1;
                for (int d = 0; d < dim; ++d) {
                    o += idx[d] * s;
                    s *= dims[d];
                }
                return o;
            }
            constexpr T& operator[](const MDIndex<dim>& idx) { return val[offset(idx)]; }
            constexpr const T& operator[](const MDIndex<dim>& idx) const { return val[offset(idx)]; }
        };
    }// namespace DS
    template <typename E, int N0, int N1, int N2>
    struct ConvExpr : ExprTag {
        E arg;
        std::array<Real, N0 * N1 * N2> ker;
        static constexpr bool has_unknown = E::has_unknown;
        static constexpr int dim = E::dim;
        mutable DS::Range<(dim > 0 ? dim : 1)> accessibleRange, localRange, logicalRange, assignableRange;
        mutable std::array<LocOnMesh, (dim > 0 ? dim : 1)> loc {};
        template <int NF, int NS>
        struct Dev {
            using D0 = typename E::template Dev<NF, NS + N0 * N1 * N2>;
            using type = opf::Conv<N0, N1, N2, NS, typename D0::type>;
            static constexpr int nf = D0::nf, ns = D0::ns;
        };
        void flatten(internal::Flat& fl) const {
            fl.sig += "Conv<" + std::to_string(N0) + "," + std::to_string(N1) + "," + std::to_string(N2) + "," + std::to_string(fl.scalars.size()) + ",";
            for (Real v : ker) fl.scalars.push_back(v);
            arg.flatten(fl);
            fl.sig += ">";
        }
        void prepare() const {
            internal::Flat fl;
            flatten(fl);
            opf_range r;
            int l[OPF_MAX_DIM];
            internal::check_rc(opf_expr_prepare(fl.sig.c_str(), fl.fields.data(), (int) fl.fields.size(), 2, &r, l), "opf_expr_prepare");
            accessibleRange = internal::from_c<dim>(r);
            internal::check_rc(opf_expr_prepare(fl.sig.c_str(), fl.fields.data(), (int) fl.fields.size(), 0, &r, nullptr), "opf_expr_prepare");
            localRange = internal::from_c<dim>(r);
            internal::check_rc(opf_expr_prepare(fl.sig.c_str(), fl.fields.data(), (int) fl.fields.size(), 3, &r, nullptr), "opf_expr_prepare");
            logicalRange = internal::from_c<dim>(r);
            assignableRange.setEmpty();
            for (int d = 0; d < dim; ++d) loc[d] = static_cast<LocOnMesh>(l[d]);
        }
    };
    template <typename E, typename D, int... ns>
    requires ExprType<E> auto conv(E&& expr, const DS::FixedSizeTensor<D, ns...>& kernel) {
        constexpr std::array<int, 3> n = [] {
            std::array<int, 3> a {1, 1, 1};
            int k = 0;
            ((a[k++] = ns), ...);
            return a;
        }();
        static_assert(sizeof...(ns) <= 3 && ((ns % 2 == 1) && ...), "conv(): odd kernel extents, at most three axes");
        using W = internal::wrapped_t<E>;
        ConvExpr<W, n[0], n[1], n[2]> c {{}, internal::wrap(std::forward<E>(expr)), {}};
        for (int i = 0; i < kernel.total; ++i) c.ker[i] = static_cast<Real>(kernel.val[i]);
        return c;
    }

    // ------------------------------------------------------------------------------------------------ kernel registration
    namespace internal {
        template <typename T>
        struct FieldExprTrait {// FieldExprTrait.hpp: the spellings generic user code reads off a field / expression type
            static constexpr int dim = std::remove_cvref_t<T>::dim;
            using elem_type = Real;
        };
        // flatten + make sure the expression's kernels are registered with the engine (once per expression type)
        template <int DIM, typename E>
        Flat flatten_and_register(const E& e) {
            Flat fl;
            e.flatten(fl);
#ifdef OPFLOW_DEVICE_KERNELS
            using DevT = typename E::template Dev<0, 0>::type;
            static const bool once = [&] {// the library's own instantiation (all dimensions) wins if it exists
                if (!opf_expr_is_registered(fl.sig.c_str()))
                    check_rc(opf_expr_register_abi(fl.sig.c_str(), &opf::launcher<DevT, 1 << (DIM - 1)>, OPF_DEVICE_ABI), "opf_expr_register");
                return true;
            }();
            (void) once;
#endif
            return fl;
        }
    }// namespace internal

    // ------------------------------------------------------------------------------------------------ split strategies
    struct ParallelPlan;
    template <typename F>
    struct AbstractSplitStrategy {// AbstractSplitStrategy.hpp:24-32
        virtual ~AbstractSplitStrategy() = default;
        virtual std::string strategyName() const = 0;
        virtual typename F::RangeType splitRange(const typename F::RangeType& range, const ParallelPlan& plan) = 0;
        virtual std::vector<typename F::RangeType> getSplitMap(const typename F::RangeType& range, const ParallelPlan& plan) = 0;
    };
    inline int getWorkerId();
    inline int getWorkerCount();
    inline ParallelPlan& getGlobalParallelPlan();

    // ------------------------------------------------------------------------------------------------ field
    struct BCInfo {// what ExprBuilder::setBC records per side (BC objects of src/Core/BC/*.hpp reduce to this on the device)
        BCType type = BCType::Undefined;
        Real value = 0;
        // FunctorDircBC / FunctorNeumBC (DircBC.hpp:83-118, NeumBC.hpp): value as a function of the (ghost or boundary) cell index,
        // type-erased over the index type; evaluated once on the host when the field is built (opf_bc_desc.face)
        std::function<Real(const int*)> functor;
        BCType getBCType() const { return type; }
    };

    template <typename D, typename M>
    struct CartesianField : ExprTag {
        static_assert(std::is_same_v<D, Real>, "the B200 engine computes in FP64 (Real = double, BasicDataTypes.hpp:30)");
        static constexpr int dim = M::dim;
        using RangeType = DS::Range<dim>;
        using IndexType = DS::MDIndex<dim>;
        using MeshType = M;
        using elem_type = D;
        static constexpr bool has_unknown = false;

        std::string name;
        M mesh;
        std::array<LocOnMesh, dim> loc {};
        std::array<DS::Pair<BCInfo>, dim> bc {};
        std::array<DS::Pair<int>, dim> ext_width {};
        int padding = 0;
        RangeType localRange, assignableRange, accessibleRange, logicalRange, storageRange;
        std::shared_ptr<AbstractSplitStrategy<CartesianField>> splitStrategy;
        opf_field_t h = nullptr;
        bool initialized = false;

        CartesianField() = default;
        CartesianField(const CartesianField& o) { copyFrom(o); }// deep copy (CartesianField.hpp:57-68)
        CartesianField(CartesianField&& o) noexcept { moveFrom(std::move(o)); }
        ~CartesianField() {
            if (h) opf_field_destroy(h);
        }

        // ---- assignment (Expr.hpp:53-117 -> assignImpl_final -> FieldAssigner::assign)
        CartesianField& operator=(const CartesianField& o) {// CartesianField.hpp:180-193
            if (this == &o) return *this;
            if (!initialized) {
                copyFrom(o);
                return *this;
            }
            o.syncToDevice();
            syncToDevice();
            internal::check_rc(opf_field_assign_field(h, OPF_OP_EQ, o.h), "opf_field_assign_field");
            touch();
            return *this;
        }
        CartesianField& operator=(CartesianField&& o) noexcept {
            if (this != &o) {
                if (h) opf_field_destroy(h);
                h = nullptr;
                moveFrom(std::move(o));
            }
            return *this;
        }
        template <typename E>
        requires(ExprType<E> && !internal::FieldType<E>) CartesianField& operator=(const E& e) { return assignExpr(OPF_OP_EQ, e); }
        template <Meta::Numerical T>
        CartesianField& operator=(T c) { return assignScalar(OPF_OP_EQ, static_cast<Real>(c)); }
#define OPF_FE_COMPOUND(sym, OPC)                                                                                      \
    template <typename E>                                                                                              \
    requires(ExprType<E> && !internal::FieldType<E>) CartesianField& operator sym(const E& e) { return assignExpr(OPC, e); } \
    CartesianField& operator sym(const CartesianField& o) {                                                            \
        o.syncToDevice();                                                                                              \
        syncToDevice();                                                                                                \
        internal::check_rc(opf_field_assign_field(h, OPC, o.h), "opf_field_assign_field");                            \
        touch();                                                                                                       \
        return *this;                                                                                                  \
    }                                                                                                                  \
    template <Meta::Numerical T>                                                                                       \
    CartesianField& operator sym(T c) { return assignScalar(OPC, static_cast<Real>(c)); }
        OPF_FE_COMPOUND(+=, OPF_OP_ADD)
        OPF_FE_COMPOUND(-=, OPF_OP_MINUS)
        OPF_FE_COMPOUND(*=, OPF_OP_MUL)
        OPF_FE_COMPOUND(/=, OPF_OP_DIV)
#undef OPF_FE_COMPOUND

        template <typename E>
        CartesianField& assignExpr(int op, const E& e) {
            requireInit("assign an expression to");
            syncToDevice();
            auto fl = internal::flatten_and_register<dim>(e);
            internal::check_rc(opf_assign(h, op, fl.sig.c_str(), fl.fields.data(), (int) fl.fields.size(), fl.scalars.data(), (int) fl.scalars.size()),
                               "opf_assign");
            touch();
            return *this;
        }
        CartesianField& assignScalar(int op, Real c) {// CartesianField.hpp:237-280
            requireInit("assign a constant to");
            syncToDevice();
            internal::check_rc(opf_field_assign_scalar(h, op, c), "opf_field_assign_scalar");
            touch();
            return *this;
        }
        // CartesianField.hpp:283-294: arbitrary host functor of the physical coordinates (x, or x + dx/2 on Center axes)
        CartesianField& initBy(const std::function<D(const std::array<Real, dim>&)>& f) {
            requireInit("initBy");
            syncToDevice();// pending host writes (operator[]) reach the device before it is overwritten, never after
            const auto w = DS::commonRange(assignableRange, localRange);
            if (w.count() > 0) {
                std::vector<Real> vals((std::size_t) w.count());
                std::size_t n = 0;
                rangeFor_s(w, [&](auto&& i) {
                    std::array<Real, dim> c;
                    for (int k = 0; k < dim; ++k) c[k] = loc[k] == LocOnMesh::Corner ? mesh.x(k, i[k]) : mesh.x(k, i[k]) + .5 * mesh.dx(k, i[k]);
                    vals[n++] = f(c);
                });
                const opf_range r = internal::to_c(w);
                internal::check_rc(opf_field_upload(h, &r, vals.data()), "opf_field_upload");
            }
            touch();
            return updatePadding();
        }
        CartesianField& updatePadding() {
            requireInit("updatePadding");
            syncToDevice();
            internal::check_rc(opf_field_update_padding(h), "opf_field_update_padding");
            touch();
            return *this;
        }
        // resplitWithStrategy (CartesianField.hpp:83-177): values kept, blocks moved between the ranks.  Like the reference's it only
        // acts on a decomposed run; the strategy itself is not stored (the reference keeps the builder's one as well).
        template <typename S>
        void resplitWithStrategy(S* strategy) {
            if (!strategy || getWorkerCount() <= 1) return;
            requireInit("resplitWithStrategy");
            syncToDevice();
            auto map = strategy->getSplitMap(mesh.getRange(), getGlobalParallelPlan());
            std::vector<opf_range> split;
            for (auto& r : map) split.push_back(internal::to_c(r));
            internal::check_rc(opf_field_resplit(h, split.data()), "opf_field_resplit");
            auto get = [&](int which) {
                opf_range r;
                internal::check_rc(opf_field_get_range(h, which, &r), "opf_field_get_range");
                return internal::from_c<dim>(r);
            };
            localRange = get(0);
            assignableRange = get(1);
            accessibleRange = get(2);
            logicalRange = get(3);
            storageRange = get(4);
            mirror.clear();
            mirror_valid = host_dirty = false;
        }
        void prepare() const {}
        const M& getMesh() const { return mesh; }
        auto getLocalWritableRange() const { return DS::commonRange(assignableRange, localRange); }
        auto getLocalReadableRange() const {
            opf_range r;
            internal::check_rc(opf_field_get_range(h, 5, &r), "opf_field_get_range");
            return internal::from_c<dim>(r);
        }
        const auto& getName() const { return name; }

        // ---- host access: a lazily synchronised mirror of the device storage (user lambdas in rangeFor / rangeReduce, writers)
        Real evalAt(const IndexType& i) const { return hostRef(i); }
        Real operator[](const IndexType& i) const { return hostRef(i); }
        Real& operator[](const IndexType& i) {
            Real& r = hostRef(i);
            host_dirty = true;
            return r;
        }
        Real operator()(const IndexType& i) const { return hostRef(i); }
        bool contains(const CartesianField& o) const { return this == &o; }
        // mirror -> device if host code wrote through operator[] (whole storage, no updatePadding: PlainTensor semantics)
        void syncToDevice() const {
            if (!host_dirty) return;
            const opf_range r = internal::to_c(storageRange);
            internal::check_rc(opf_field_upload(h, &r, mirror.data()), "opf_field_upload");
            host_dirty = false;
        }
        void touch() const { mirror_valid = false; }// device data changed

    private:
        template <typename>
        friend struct ExprBuilder;
        mutable std::vector<Real> mirror;
        mutable bool mirror_valid = false, host_dirty = false;

        void requireInit(const char* what) const {
            if (!initialized || !h) {
                OP_CRITICAL("CartesianField not initialized. Cannot {} it.", what);
                OP_ABORT;
            }
        }
        Real& hostRef(const IndexType& i) const {
            requireInit("read");
            if (!mirror_valid) {
                mirror.resize((std::size_t) storageRange.count());
                const opf_range r = internal::to_c(storageRange);
                internal::check_rc(opf_field_download(h, &r, mirror.data()), "opf_field_download");
                mirror_valid = true;
            }
            std::size_t off = 0, stride = 1;
            for (int d = 0; d < dim; ++d) {
                off += (std::size_t)(i[d] - storageRange.start[d]) * stride;
                stride *= (std::size_t)(storageRange.end[d] - storageRange.start[d]);
            }
            return mirror[off];
        }
        void copyMeta(const CartesianField& o) {
            name = o.name;
            mesh = o.mesh;
            loc = o.loc;
            bc = o.bc;
            ext_width = o.ext_width;
            padding = o.padding;
            localRange = o.localRange;
            assignableRange = o.assignableRange;
            accessibleRange = o.accessibleRange;
            logicalRange = o.logicalRange;
            storageRange = o.storageRange;
            splitStrategy = o.splitStrategy;
            initialized = o.initialized;
        }
        void copyFrom(const CartesianField& o) {
            if (h) opf_field_destroy(h);
            h = nullptr;
            copyMeta(o);
            mirror_valid = host_dirty = false;
            if (o.h) {
                o.syncToDevice();
                h = internal::check_ptr(opf_field_clone(o.h, name.c_str()), "opf_field_clone");
            }
        }
        void moveFrom(CartesianField&& o) {
            copyMeta(o);
            h = o.h;
            o.h = nullptr;
            o.initialized = false;
            mirror = std::move(o.mirror);
            mirror_valid = o.mirror_valid;
            host_dirty = o.host_dirty;
        }
    };

    // ExprBuilder<CartesianField> (CartesianField.hpp:796-1033).  Like the reference, a builder keeps its settings between
    // build() calls and build() hands out a reference to its internal field (callers copy-construct from it).
    template <typename D, typename M>
    struct ExprBuilder<CartesianField<D, M>> {
        using Field = CartesianField<D, M>;
        static constexpr int dim = Field::dim;
        Field f;

        ExprBuilder() = default;
        auto& setName(const std::string& n) {
            f.name = n;
            return *this;
        }
        auto& setMesh(const M& m) {
            f.mesh = m;
            return *this;
        }
        auto& setLoc(const std::array<LocOnMesh, dim>& l) {
            f.loc = l;
            return *this;
        }
        auto& setLoc(LocOnMesh l) {
            f.loc.fill(l);
            return *this;
        }
        auto& setLocOfDim(int i, LocOnMesh l) {
            f.loc[i] = l;
            return *this;
        }
        // setBC(d, pos, type) for logical BCs, setBC(d, pos, type, value) for Dirc / Neum (CartesianField.hpp:827-895)
        auto& setBC(int d, DimPos pos, BCType type) {// only the requested side, like the reference (CartesianField.hpp:827-850)
            (pos == DimPos::start ? f.bc[d].start : f.bc[d].end) = BCInfo {type, 0.};
            return *this;
        }
        template <Meta::Numerical T>
        auto& setBC(int d, DimPos pos, BCType type, T val) {
            // logical types forward to the value-less overload (the reference's own tests use this form,
            // CSRMatrixGeneratorMPITest.cpp:389)
            if (type == BCType::Periodic || type == BCType::Symm || type == BCType::ASymm) return setBC(d, pos, type);
            if (type != BCType::Dirc && type != BCType::Neum) {
                OP_ERROR("BC type not supported.");
                OP_ABORT;
            }
            (pos == DimPos::start ? f.bc[d].start : f.bc[d].end) = BCInfo {type, static_cast<Real>(val)};
            return *this;
        }
        // functor BC: setBC(d, pos, BCType::Dirc | Neum, [](auto&& index) { ... }) (CartesianField.hpp:870-892)
        template <typename Fn>
        requires(!Meta::Numerical<Fn> && requires(Fn fn, DS::MDIndex<dim> i) {
            { fn(i) } -> std::convertible_to<Real>;
        }) auto& setBC(int d, DimPos pos, BCType type, Fn&& functor) {
            if (type != BCType::Dirc && type != BCType::Neum) {
                OP_ERROR("BC Type not supported.");
                OP_ABORT;
            }
            BCInfo info {type, 0.};
            info.functor = [fn = std::decay_t<Fn>(std::forward<Fn>(functor))](const int* idx) {
                DS::MDIndex<dim> i;
                for (int k = 0; k < dim; ++k) i[k] = idx[k];
                return static_cast<Real>(fn(i));
            };
            (pos == DimPos::start ? f.bc[d].start : f.bc[d].end) = std::move(info);
            return *this;
        }
        auto& setExt(int d, DimPos pos, int width) {
            (pos == DimPos::start ? f.ext_width[d].start : f.ext_width[d].end) = width;
            return *this;
        }
        auto& setExt(int width) {
            for (auto& e : f.ext_width) e.start = e.end = width;
            return *this;
        }
        auto& setPadding(int p) {
            f.padding = p;
            return *this;
        }
        auto& setSplitStrategy(std::shared_ptr<AbstractSplitStrategy<Field>> s) {
            f.splitStrategy = std::move(s);
            return *this;
        }

        auto& build() {// calculateRanges + validateRanges + storage + updatePadding happen in opf_field_create
            if (!f.mesh.h()) {
                OP_CRITICAL("ExprBuilder: setMesh() missing for field '{}'", f.name);
                OP_ABORT;
            }
            opf_field_desc d {};
            d.mesh = f.mesh.h();
            for (int k = 0; k < dim; ++k) {
                d.loc[k] = static_cast<int>(f.loc[k]);
                d.bc[k][0].type = static_cast<int>(f.bc[k].start.type);
                d.bc[k][0].value = f.bc[k].start.value;
                d.bc[k][1].type = static_cast<int>(f.bc[k].end.type);
                d.bc[k][1].value = f.bc[k].end.value;
                d.ext[k][0] = f.ext_width[k].start;
                d.ext[k][1] = f.ext_width[k].end;
            }
            d.padding = f.padding;
            std::vector<opf_range> split;
            if (f.splitStrategy && getWorkerCount() > 1) {
                auto map = f.splitStrategy->getSplitMap(f.mesh.getRange(), getGlobalParallelPlan());
                for (auto& r : map) split.push_back(internal::to_c(r));
                d.n_ranks = (int) split.size();
                d.rank = getWorkerId();
                d.split_map = split.data();
            }
            // functor BCs: one value per index of the face slab -- the ghost layers of that side plus the boundary node / cell row,
            // over the whole logical extent of the other axes (where updatePadding evaluates bc->evalAt(index),
            // CartesianField.hpp:351-606).  The slab comes from a device-free plan of the same description.
            std::vector<std::vector<double>> faces;
            bool any_functor = false;
            for (int k = 0; k < dim; ++k) any_functor = any_functor || f.bc[k].start.functor || f.bc[k].end.functor;
            if (any_functor) {
                opf_field_t plan = internal::check_ptr(opf_field_plan(&d, f.name.c_str()), "opf_field_plan");
                opf_range logical, acc;
                internal::check_rc(opf_field_get_range(plan, 3, &logical), "opf_field_get_range");
                internal::check_rc(opf_field_get_range(plan, 2, &acc), "opf_field_get_range");
                opf_field_destroy(plan);
                faces.reserve(2 * dim);
                for (int k = 0; k < dim; ++k)
                    for (int side = 0; side < 2; ++side) {
                        const BCInfo& b = side == 0 ? f.bc[k].start : f.bc[k].end;
                        if (!b.functor) continue;
                        opf_range fr = logical;
                        for (int a = dim; a < OPF_MAX_DIM; ++a) fr.start[a] = 0, fr.end[a] = 1;
                        if (side == 0) fr.end[k] = acc.start[k] + 1;
                        else
                            fr.start[k] = acc.end[k] - 1;
                        std::vector<double> vals;
                        int idx[OPF_MAX_DIM] = {0, 0, 0};
                        for (idx[2] = fr.start[2]; idx[2] < fr.end[2]; ++idx[2])
                            for (idx[1] = fr.start[1]; idx[1] < fr.end[1]; ++idx[1])
                                for (idx[0] = fr.start[0]; idx[0] < fr.end[0]; ++idx[0]) vals.push_back(b.functor(idx));
                        faces.push_back(std::move(vals));
                        d.bc[k][side].face = faces.back().data();
                        d.bc[k][side].face_range = fr;
                    }
            }
            if (f.h) opf_field_destroy(f.h);
            f.h = internal::check_ptr(opf_field_create(&d, f.name.c_str()), "opf_field_create");
            auto get = [&](int which) {
                opf_range r;
                internal::check_rc(opf_field_get_range(f.h, which, &r), "opf_field_get_range");
                return internal::from_c<dim>(r);
            };
            f.localRange = get(0);
            f.assignableRange = get(1);
            f.accessibleRange = get(2);
            f.logicalRange = get(3);
            f.storageRange = get(4);
            f.padding = opf_field_padding(f.h);
            f.initialized = true;
            f.mirror_valid = f.host_dirty = false;
            return f;
        }
    };
}// namespace OpFlow
