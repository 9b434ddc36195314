// opflow/base.hpp -- basic types of the B200 front-end: the spellings user programs touch, re-expressed
// in C++20 so that nvcc 12.9 compiles them.  Reference: src/Core/BasicDataTypes.hpp, src/Core/Constants.hpp, src/Core/Meta.hpp,
// src/Core/Macros.hpp, src/DataStructures/Index/MDIndex.hpp, src/DataStructures/Range/Ranges.hpp, src/DataStructures/Pair.hpp.
#pragma once
#include <opflow_b200.h>

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <format>
#include <fstream>
#include <functional>
#include <thread>
#include <iostream>
#include <memory>
#include <numeric>
#include <optional>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

namespace OpFlow {
    using Real = double;// BasicDataTypes.hpp:27-31 (Real is always double in the reference build)
    using Index = int;
    using Size = std::size_t;
    inline constexpr Real PI = 3.14159265358979323846;// Constants.hpp

    namespace Meta {
        template <int N>
        struct int_ {
            static constexpr int value = N;
        };
        template <typename T>
        using RealType = std::remove_cvref_t<T>;
        template <typename T>
        concept Numerical = std::is_arithmetic_v<std::remove_cvref_t<T>>;
    }// namespace Meta

    namespace Math {
        inline constexpr auto pow2(double a) { return a * a; }
        inline constexpr auto pow3(double a) { return a * a * a; }
        inline auto norm2(double a, double b) { return std::sqrt(a * a + b * b); }
        inline auto norm2(double a, double b, double c) { return std::sqrt(a * a + b * b + c * c); }
        inline auto mid(double a, double b) { return (a + b) * 0.5; }
        inline int int_pow(int b, int e) {
            int r = 1;
            for (int i = 0; i < e; ++i) r *= b;
            return r;
        }
#if defined(__CUDACC__)
#define OPF_HD __host__ __device__
#else
#define OPF_HD
#endif
        // Math/Function/Numeric.hpp:86-96; usable inside device functors (UniOpAdaptor of examples/LevelSet/UniLS.cpp:109-112)
        OPF_HD inline double smoothDelta(double eps, double d) {
            return (d > -d ? d : -d) > eps ? 0. : 1. / 2. / eps * (1 + std::cos(d * 3.141592653589793 / eps));
        }
        OPF_HD inline double smoothHeviside(double eps, double d) {
            if (d < -eps) return 0.;
            if (d > eps) return 1.;
            return 0.5 * (1. + d / eps + 1. / PI * std::sin(PI * d / eps));
        }
    }// namespace Math

    // Constants.hpp:48-53, BC/BCBase.hpp:24 (numeric values equal the C ABI enums)
    enum class DimPos { start, end };
    enum class LocOnMesh { Corner, Center };
    enum class BCType { Undefined, Dirc, Neum, Periodic, Internal, Symm, ASymm };
    enum class BasicArithOp { Eq, Add, Minus, Mul, Div, Mod, And, Or, Xor, LShift, RShift };
    enum class MeshExtMode { Undefined, Symm, Periodic, Uniform };
    inline bool isLogicalBC(BCType t) { return t == BCType::Periodic || t == BCType::Internal || t == BCType::Symm || t == BCType::ASymm; }

    // ---- logging / abort macros (Macros.hpp:66-266).  The reference's logging compiles to nothing at run time; here it prints.
    namespace internal {
        template <typename... A>
        void log_line(const char* level, std::format_string<A...> f, A&&... a) {
            std::fprintf(stderr, "[%s] %s\n", level, std::format(f, std::forward<A>(a)...).c_str());
        }
        inline void check_rc(int rc, const char* what) {
            if (rc != 0) {
                std::fprintf(stderr, "[critical] %s failed (%d): %s\n", what, rc, opf_last_error());
                std::exit(1);// OP_ABORT == std::exit(1) (Macros.hpp:146-151)
            }
        }
        template <typename T>
        T* check_ptr(T* p, const char* what) {
            if (!p) {
                std::fprintf(stderr, "[critical] %s failed: %s\n", what, opf_last_error());
                std::exit(1);
            }
            return p;
        }
    }// namespace internal
#define OP_INFO(...) ::OpFlow::internal::log_line("info", __VA_ARGS__)
#define OP_WARN(...) ::OpFlow::internal::log_line("warn", __VA_ARGS__)
#define OP_ERROR(...) ::OpFlow::internal::log_line("error", __VA_ARGS__)
#define OP_CRITICAL(...) ::OpFlow::internal::log_line("critical", __VA_ARGS__)
#define OP_DEBUG(...) ((void) 0)
#define OP_TRACE(...) ((void) 0)
#define OP_MPI_MASTER_INFO(...)                                                                                        \
    do {                                                                                                               \
        if (::opf_comm_rank() == 0) ::OpFlow::internal::log_line("info", __VA_ARGS__);                                 \
    } while (0)
#define OP_ABORT std::exit(1)
#define OP_NOT_IMPLEMENTED                                                                                             \
    do {                                                                                                               \
        OP_CRITICAL("not implemented: {}:{}", __FILE__, __LINE__);                                                     \
        OP_ABORT;                                                                                                      \
    } while (0)
#define OP_ASSERT(x) ((void) 0)
#define OP_ASSERT_MSG(x, ...) ((void) 0)
#define OP_PERFECT_FOWD(x) std::forward<decltype(x)>(x)

    namespace DS {
        template <typename T>
        struct Pair {
            T start {}, end {};
        };

        // MDIndex<d> (MDIndex.hpp:32-150): a POD here (the reference's has a vtable through StringifiableObj)
        template <std::size_t d>
        struct MDIndex {
            static constexpr auto dim = d;
            std::array<int, d> idx {};
            constexpr MDIndex() = default;
            constexpr explicit MDIndex(const std::array<int, d>& a) : idx(a) {}
            template <typename... T>
            requires(sizeof...(T) == d && d >= 1 && (std::is_integral_v<std::remove_cvref_t<T>> && ...)) constexpr MDIndex(T... i)
                : idx {static_cast<int>(i)...} {}
            constexpr const int& operator[](int i) const { return idx[i]; }
            constexpr int& operator[](int i) { return idx[i]; }
            constexpr const auto& get() const { return idx; }
            constexpr void set(const std::array<int, d>& o) { idx = o; }
            constexpr bool operator==(const MDIndex& o) const { return idx == o.idx; }
            constexpr MDIndex operator+(const MDIndex& o) const {
                MDIndex r = *this;
                for (std::size_t i = 0; i < d; ++i) r.idx[i] += o.idx[i];
                return r;
            }
            constexpr MDIndex operator-(const MDIndex& o) const {
                MDIndex r = *this;
                for (std::size_t i = 0; i < d; ++i) r.idx[i] -= o.idx[i];
                return r;
            }
            template <std::size_t k>
            constexpr MDIndex next(int steps = 1) const {
                MDIndex r = *this;
                r.idx[k] += steps;
                return r;
            }
            template <std::size_t k>
            constexpr MDIndex prev(int steps = 1) const {
                MDIndex r = *this;
                r.idx[k] -= steps;
                return r;
            }
            std::string toString() const {
                std::string s = "{";
                for (std::size_t i = 0; i < d; ++i) s += (i ? ", " : "") + std::to_string(idx[i]);
                return s + "}";
            }
        };

        // Range<d> (Ranges.hpp:33-225): half-open box with stride (always 1 on the device path)
        template <std::size_t d>
        struct Range {
            static constexpr int dim = d;
            using base_index_type = MDIndex<d>;
            std::array<int, d> start {}, end {}, stride;
            constexpr Range() { stride.fill(1); }
            constexpr explicit Range(const std::array<int, d>& e) : end(e) { stride.fill(1); }
            constexpr Range(const std::array<int, d>& s, const std::array<int, d>& e) : start(s), end(e) { stride.fill(1); }
            static constexpr Range EmptyRange() { return Range(); }
            constexpr int count() const {// int like the reference (Ranges.hpp:90-99)
                int c = 1;
                for (std::size_t i = 0; i < d; ++i) {
                    const int p = (end[i] - start[i] + stride[i] - 1) / stride[i];
                    if (p <= 0) return 0;
                    c *= p;
                }
                return c;
            }
            constexpr bool empty() const { return count() <= 0; }
            constexpr auto getExtends() const {
                std::array<int, d> r {};
                for (std::size_t i = 0; i < d; ++i) r[i] = end[i] - start[i];
                return r;
            }
            constexpr auto getOffset() const { return start; }
            constexpr Range slice(std::size_t k, int pos) const {
                Range r = *this;
                r.start[k] = pos;
                r.end[k] = pos + 1;
                return r;
            }
            constexpr Range slice(std::size_t k, int s, int e) const {
                Range r = *this;
                r.start[k] = s;
                r.end[k] = e;
                return r;
            }
            constexpr Range getInnerRange(int w) const {
                Range r = *this;
                for (std::size_t i = 0; i < d; ++i) {
                    r.start[i] += w;
                    r.end[i] -= w;
                }
                return r;
            }
            void setEmpty() { *this = EmptyRange(); }
            auto first() const { return base_index_type {start}; }
            auto last() const {
                auto r = base_index_type {end};
                for (std::size_t i = 0; i < d; ++i) r[i]--;
                return r;
            }
            auto center() const {
                auto r = base_index_type {start};
                for (std::size_t i = 0; i < d; ++i) r[i] = (r[i] + end[i]) / 2;
                return r;
            }
            constexpr bool operator==(const Range& o) const { return start == o.start && end == o.end && stride == o.stride; }
            std::string toString() const { return std::format("{{{} - {}}}", base_index_type {start}.toString(), base_index_type {end}.toString()); }
        };
        template <std::size_t d>
        constexpr Range<d> commonRange(const Range<d>& a, const Range<d>& b) {// Ranges.hpp:232-253
            Range<d> r;
            for (std::size_t i = 0; i < d; ++i) {
                r.start[i] = std::max(a.start[i], b.start[i]);
                r.end[i] = std::min(a.end[i], b.end[i]);
            }
            return r;
        }
        template <std::size_t d, typename T>
        constexpr bool inRange(const Range<d>& r, const T& t) {
            bool ret = true;
            for (std::size_t i = 0; i < d; ++i) ret &= (r.start[i] <= t[i] && t[i] < r.end[i]);
            return ret;
        }
    }// namespace DS

    namespace internal {
        template <std::size_t d>
        opf_range to_c(const DS::Range<d>& r) {
            opf_range o;
            for (int i = 0; i < OPF_MAX_DIM; ++i) {
                o.start[i] = i < (int) d ? r.start[i] : 0;
                o.end[i] = i < (int) d ? r.end[i] : 1;
            }
            return o;
        }
        template <std::size_t d>
        DS::Range<d> from_c(const opf_range& r) {
            DS::Range<d> o;
            for (std::size_t i = 0; i < d; ++i) {
                o.start[i] = r.start[i];
                o.end[i] = r.end[i];
            }
            return o;
        }
    }// namespace internal

    // ---- execution engine, host side (RangeFor.hpp:39-135, StructFor.hpp).  Arbitrary host lambdas stay on the host (they
    // may capture host state, RangeForTest.cpp:131-175); device expressions go through Field assignment / rangeReduce(expr).
    template <std::size_t d, typename F>
    F rangeFor_s(const DS::Range<d>& range, F&& func) {
        const int total = range.count();
        if (total <= 0) return std::forward<F>(func);
        DS::MDIndex<d> i {range.start};
        for (int c = 0; c < total; ++c) {
            func(i);
            for (std::size_t k = 0; k < d; ++k) {// RangedIndex::operator++ carry chain (RangedIndex.hpp:158-179), axis 0 fastest
                i[k] += range.stride[k];
                if (i[k] < range.end[k] || k == d - 1) break;
                i[k] = range.start[k];
            }
        }
        return std::forward<F>(func);
    }
    namespace internal {
        // shared-memory workers of the global plan (setGlobalParallelPlan, ParallelPlan.hpp:38-50); 1 by default, like the reference
        inline int g_host_threads = 1;
    }
    // rangeFor (RangeFor.hpp:69-84): the reference hands the range to tbb::parallel_for with the plan's shared-memory worker count
    // (1 unless the program asks for more).  Here: the slowest axis is cut into one contiguous piece per worker.  The first index runs
    // on the calling thread before the workers start, so that lazily synchronised host mirrors of the fields the functor touches are
    // in place (operator[] downloads on first access); like the reference, the functor must be safe to call concurrently.
    template <std::size_t d, typename F>
    F rangeFor(const DS::Range<d>& range, F&& func) {
        const int nt = internal::g_host_threads;
        const int total = range.count();
        const int n_slow = range.end[d - 1] - range.start[d - 1];
        if (nt <= 1 || total < 4096 || n_slow < 2 || range.stride[d - 1] != 1) return rangeFor_s(range, std::forward<F>(func));
        {
            DS::MDIndex<d> first {range.start};
            func(first);
        }
        const int workers = std::min(nt, n_slow);
        std::vector<std::thread> pool;
        pool.reserve(workers);
        for (int w = 0; w < workers; ++w) {
            DS::Range<d> piece = range;
            piece.start[d - 1] = range.start[d - 1] + (int) ((long long) n_slow * w / workers);
            piece.end[d - 1] = range.start[d - 1] + (int) ((long long) n_slow * (w + 1) / workers);
            pool.emplace_back([piece, w, &func, &range] {
                bool skip_first = w == 0;// already done above
                rangeFor_s(piece, [&](auto&& i) {
                    if (skip_first) {
                        skip_first = false;
                        return;
                    }
                    func(i);
                });
            });
        }
        for (auto& t : pool) t.join();
        return std::forward<F>(func);
    }
    template <std::size_t d, typename ReOp, typename F>
    auto rangeReduce_s(const DS::Range<d>& range, ReOp&& op, F&& func) {
        using R = std::remove_cvref_t<decltype(func(std::declval<DS::MDIndex<d>&>()))>;
        R acc {};
        rangeFor_s(range, [&](auto&& i) { acc = op(acc, func(i)); });
        return acc;
    }
    template <std::size_t d, typename ReOp, typename F>
    auto rangeReduce(const DS::Range<d>& range, ReOp&& op, F&& func) {
        return rangeReduce_s(range, std::forward<ReOp>(op), std::forward<F>(func));
    }
}// namespace OpFlow
