// opflow/solve.hpp -- the implicit path of the B200 front-end: `lhs == rhs` equations, StructSolverParams, PrecondStructSolver,
// EqnSolveHandler / makeEqnSolveHandler / Solve.
// Reference: src/Core/Equation/Equation.hpp, EqnSolveHandler.hpp:17-33, HYPREEqnSolveHandler.hpp:44-231, UnifiedSolve.hpp:36-43,
// src/Core/Solvers/Struct/StructSolver.hpp:23-51 and StructSolver{PCG,GMRES,BiCGSTAB,PFMG,SMG,Jacobi,...}.hpp (parameter structs),
// StructSolverPrecond.hpp:19-71.
// The handler never assembles a matrix: it hands the equation's two sides to opf_solver_* as device expressions; the unknown `e`
// of the equation lambda is a leaf flagged in the unknown mask.
#pragma once
#include "field.hpp"

namespace OpFlow {
    template <typename L, typename R>
    struct Equation {
        L lhs;
        R rhs;
    };
    // `lhs == rhs` between expressions builds an equation (Equation.hpp) -- point-wise equality is spelt conditional(a - b, ...)
    template <typename A, typename B>
    requires internal::ExprOperands<A, B> auto operator==(A&& a, B&& b) {
        return Equation<internal::wrapped_t<A>, internal::wrapped_t<B>> {internal::wrap(std::forward<A>(a)), internal::wrap(std::forward<B>(b))};
    }

    enum class StructSolverType { None, Jacobi, SMG, PFMG, CYCRED, PCG, GMRES, FGMRES, LGMRES, BICGSTAB };// StructSolver.hpp:23-34

    struct StructSolverParamsBase {// StructSolver.hpp:39-51
        std::optional<Real> tol {};
        std::optional<int> maxIter {};
        int comm = 0;
        bool staticMat = false;
        bool pinValue = false;
        std::optional<std::string> dumpPath {};
    };
    template <StructSolverType type>
    struct StructSolverParams;
    template <>
    struct StructSolverParams<StructSolverType::None> : StructSolverParamsBase {};
    template <>
    struct StructSolverParams<StructSolverType::Jacobi> : StructSolverParamsBase {
        std::optional<bool> useZeroGuess;
    };
    template <>
    struct StructSolverParams<StructSolverType::PFMG> : StructSolverParamsBase {// StructSolverPFMG.hpp:24-33
        std::optional<int> maxLevels {}, relChange {};
        std::optional<bool> useZeroGuess {};
        std::optional<int> relaxType {};
        std::optional<Real> jacobiWeight {};
        std::optional<int> rapType {};
        std::optional<int> numPreRelax {}, numPostRelax {}, skipRelax {};
        std::optional<std::vector<Real>> dxyz {};
        std::optional<int> logging {}, printLevel {};
    };
    template <>
    struct StructSolverParams<StructSolverType::SMG> : StructSolverParamsBase {
        std::optional<int> memoryUse {}, relChange {};
        std::optional<bool> useZeroGuess {};
        std::optional<int> numPreRelax {}, numPostRelax {};
        std::optional<int> logging {}, printLevel {};
    };
    template <>
    struct StructSolverParams<StructSolverType::CYCRED> : StructSolverParamsBase {
        std::optional<int> tDim {};
    };
    template <>
    struct StructSolverParams<StructSolverType::PCG> : StructSolverParamsBase {// StructSolverPCG.hpp:20-23
        std::optional<Real> absTol;
        std::optional<int> twoNorm, relChange, logging, printLevel;
    };
    template <>
    struct StructSolverParams<StructSolverType::GMRES> : StructSolverParamsBase {// StructSolverGMRES.hpp:20-23
        std::optional<Real> absTol;
        std::optional<int> kDim, logging, printLevel;
    };
    template <>
    struct StructSolverParams<StructSolverType::FGMRES> : StructSolverParamsBase {
        std::optional<Real> absTol;
        std::optional<int> kDim, logging, printLevel;
    };
    template <>
    struct StructSolverParams<StructSolverType::LGMRES> : StructSolverParamsBase {
        std::optional<Real> absTol;
        std::optional<int> kDim, augDim, logging, printLevel;
    };
    template <>
    struct StructSolverParams<StructSolverType::BICGSTAB> : StructSolverParamsBase {// StructSolverBiCGSTAB.hpp:20-23
        std::optional<Real> absTol;
        std::optional<int> logging, printLevel;
    };

    template <StructSolverType Type>
    struct StructSolver {
        static constexpr StructSolverType type = Type, precType = StructSolverType::None;
        StructSolverParams<Type> params;
        StructSolverParams<StructSolverType::None> precParams;
        StructSolver() = default;
        explicit StructSolver(const StructSolverParams<Type>& p) : params(p) {}
    };
    template <StructSolverType Type, StructSolverType PType>
    struct PrecondStructSolver {// StructSolverPrecond.hpp:19-71
        static constexpr StructSolverType type = Type, precType = PType;
        StructSolverParams<Type> params;
        StructSolverParams<PType> precParams;
        PrecondStructSolver() = default;
        PrecondStructSolver(const StructSolverParams<Type>& p, const StructSolverParams<PType>& pp) : params(p), precParams(pp) {}
    };

    struct EqnSolveState {// EqnSolveHandler.hpp:17-25
        int niter = 0;
        double relerr = 0, abserr = 0;
        EqnSolveState() = default;
        explicit EqnSolveState(int n) : niter(n) {}
        explicit EqnSolveState(double e) : relerr(e) {}
        EqnSolveState(int n, double e) : niter(n), relerr(e) {}
        EqnSolveState(int n, double e, double a) : niter(n), relerr(e), abserr(a) {}
    };
    struct EqnSolveHandler {// EqnSolveHandler.hpp:28-33
        virtual ~EqnSolveHandler() = default;
        virtual void init() = 0;
        virtual EqnSolveState solve() = 0;
        virtual void generateAb() {}
    };

    namespace internal {
        template <typename P>
        void fill_precond_params(opf_solver_params& o, const P& pp) {
            if constexpr (requires { pp.numPreRelax; }) {
                if (pp.numPreRelax) o.num_pre_relax = *pp.numPreRelax;
                if (pp.numPostRelax) o.num_post_relax = *pp.numPostRelax;
            }
            if constexpr (requires { pp.relaxType; })
                if (pp.relaxType) o.relax_type = *pp.relaxType;
            if (pp.tol) o.precond_tol = *pp.tol;
            if (pp.maxIter) o.precond_max_iter = *pp.maxIter;
        }
    }// namespace internal

    // HYPREEqnSolveHandler's role (HYPREEqnSolveHandler.hpp:50-231) on the matrix-free engine
    template <typename F, typename T, typename S>
    struct GpuEqnSolveHandler : EqnSolveHandler {
        F eqn_getter;
        T* target;
        S solver;
        opf_solver_t h = nullptr;
        std::string lhs_sig;

        GpuEqnSolveHandler(const F& f, T& t, const S& s) : eqn_getter(f), target(&t), solver(s) {}
        ~GpuEqnSolveHandler() override {
            if (h) opf_solver_destroy(h);
        }
        void init() override {}

        opf_solver_params makeParams() const {
            opf_solver_params p {};
            p.type = static_cast<int>(S::type);
            p.precond = static_cast<int>(S::precType);
            p.tol = solver.params.tol.value_or(0.);
            p.max_iter = solver.params.maxIter.value_or(0);
            p.static_mat = solver.params.staticMat;
            p.pin_value = solver.params.pinValue;
            p.precond_max_iter = 1;
            p.num_pre_relax = p.num_post_relax = 1;
            p.relax_type = 1;
            if constexpr (requires { solver.params.kDim; }) {
                if (solver.params.kDim) p.k_dim = *solver.params.kDim;
                if constexpr (requires { solver.params.augDim; })
                    if (solver.params.augDim) p.k_dim = (p.k_dim > 0 ? p.k_dim : 5) + *solver.params.augDim;
            }
            if constexpr (S::precType != StructSolverType::None) internal::fill_precond_params(p, solver.precParams);
            else
                internal::fill_precond_params(p, solver.params);
            if (S::precType != StructSolverType::None) p.precond_tol = 0.;// HYPRE runs the preconditioner for a fixed number of cycles
            if (p.precond_max_iter <= 0 || S::precType != StructSolverType::None) p.precond_max_iter = 1;
            return p;
        }

        template <typename LHS, typename RHS>
        EqnSolveState run(const LHS& lhs, const RHS& rhs) {
            target->syncToDevice();
            auto fl = internal::flatten_and_register<T::dim>(lhs);
#ifdef OPFLOW_DEVICE_KERNELS
            {// fused residual r = b - lhs(x): the same tree with its field leaves shifted by one behind a leading F<0>
                using ResT = opf::Sub<opf::F<0>, typename LHS::template Dev<1, 0>::type>;
                static const bool once = [&] {
                    internal::Flat sh;
                    sh.fields.push_back(nullptr);
                    lhs.flatten(sh);
                    const std::string rs = "Sub<F<0>," + sh.sig + ">";
                    if (!opf_expr_is_registered(rs.c_str())) internal::check_rc(opf_expr_register_abi(rs.c_str(), &opf::launcher<ResT, 1 << (T::dim - 1)>, OPF_DEVICE_ABI), "opf_expr_register");
                    return true;
                }();
                (void) once;
            }
#endif
            if (!h) {
                const opf_solver_params p = makeParams();
                lhs_sig = fl.sig;
                h = internal::check_ptr(opf_solver_create(target->h, fl.sig.c_str(), fl.fields.data(), (int) fl.fields.size(), fl.scalars.data(),
                                                          (int) fl.scalars.size(), fl.mask, &p),
                                        "opf_solver_create");
            } else {
                internal::check_rc(opf_solver_update(h, fl.fields.data(), (int) fl.fields.size(), fl.scalars.data(), (int) fl.scalars.size()), "opf_solver_update");
            }
            auto fr = internal::flatten_and_register<T::dim>(rhs);
            opf_solve_state st {};
            internal::check_rc(opf_solver_solve(h, fr.sig.c_str(), fr.fields.data(), (int) fr.fields.size(), fr.scalars.data(), (int) fr.scalars.size(), &st),
                               "opf_solver_solve");
            target->touch();
            return EqnSolveState {st.niter, st.relerr, st.abserr};
        }

        EqnSolveState solve() override {
            auto eq = eqn_getter(UnknownRef<T> {{}, target});
            using L = std::remove_cvref_t<decltype(eq.lhs)>;
            using R = std::remove_cvref_t<decltype(eq.rhs)>;
            static_assert(L::has_unknown || R::has_unknown, "the equation does not contain its unknown");
            if constexpr (!R::has_unknown) return run(eq.lhs, eq.rhs);// lhs(e) == rhs          (Poisson form: multigrid-capable)
            else if constexpr (!L::has_unknown)
                return run(eq.rhs, eq.lhs);
            else// e on both sides: solve (lhs - rhs)(e) == 0; the engine splits off the e-free part numerically (affine operator)
                return run(makeExpression<SubOp>(eq.lhs, eq.rhs), ScalarExpr {{}, 0.});
        }
    };

    // makeEqnSolveHandler(f, target, solver) (HYPREEqnSolveHandler.hpp:44-48)
    template <typename F, typename T, typename S>
    requires internal::FieldType<T> std::unique_ptr<EqnSolveHandler> makeEqnSolveHandler(F&& f, T&& target, S&& solver) {
        using H = GpuEqnSolveHandler<std::remove_cvref_t<F>, std::remove_cvref_t<T>, std::remove_cvref_t<S>>;
        return std::make_unique<H>(f, target, solver);
    }

    // Solve<type, pType>(func, target, params, precParams) (UnifiedSolve.hpp:36-43)
    template <StructSolverType type = StructSolverType::GMRES, StructSolverType pType = StructSolverType::None, typename F, typename T>
    requires internal::FieldType<T> auto Solve(const F& func, T&& target, StructSolverParams<type> params = StructSolverParams<type> {},
                                               StructSolverParams<pType> precParams = StructSolverParams<pType> {}) {
        auto solver = PrecondStructSolver<type, pType>(params, precParams);
        auto handler = makeEqnSolveHandler(func, target, solver);
        return handler->solve();
    }
}// namespace OpFlow
