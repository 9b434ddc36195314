// opflow/mesh.hpp -- CartesianMesh<Meta::int_<N>> and MeshBuilder over the engine's opf_mesh_* entry points.
// Reference: src/Core/Mesh/Structured/CartesianMesh.hpp:35-304 (mesh :35-120, builder :122-304).  The coordinate arithmetic
// (x[k] = (max-min)/(n-1)*(i-start)+min, extension by 5 cells, x integrated outward from dx) is done once by the engine in the
// reference's operation order (engine_core.cu) and mirrored back here, so host accessors return bit-identical numbers.
#pragma once
#include "base.hpp"

namespace OpFlow {
    template <typename Dim>
    struct CartesianMesh;

    namespace internal {
        template <typename M>
        struct CartesianMeshTrait;
        template <int N>
        struct CartesianMeshTrait<CartesianMesh<Meta::int_<N>>> {
            static constexpr int dim = N;
        };
        struct MeshHandle {// shared ownership of the engine object (fields hold their own engine-side reference)
            opf_mesh_t h = nullptr;
            explicit MeshHandle(opf_mesh_t m) : h(m) {}
            MeshHandle(const MeshHandle&) = delete;
            ~MeshHandle() {
                if (h) opf_mesh_destroy(h);
            }
        };
    }// namespace internal

    template <int N>
    struct CartesianMesh<Meta::int_<N>> {
        static constexpr int dim = N;
        std::shared_ptr<internal::MeshHandle> handle;
        std::array<int, N> dims {};
        DS::Range<N> range, extRange;
        std::array<std::vector<Real>, N> _x, _dx, _idx;// index 0 <-> extRange.start (CartesianMesh.hpp:39-49)

        CartesianMesh() = default;
        opf_mesh_t h() const { return handle ? handle->h : nullptr; }
        auto x(int d, int i) const { return _x[d][i - extRange.start[d]]; }
        auto dx(int d, int i) const { return _dx[d][i - extRange.start[d]]; }
        auto idx(int d, int i) const { return _idx[d][i - extRange.start[d]]; }
        template <std::size_t dd>
        auto x(int d, const DS::MDIndex<dd>& i) const {
            return x(d, i[d]);
        }
        template <std::size_t dd>
        auto dx(int d, const DS::MDIndex<dd>& i) const {
            return dx(d, i[d]);
        }
        template <std::size_t dd>
        auto idx(int d, const DS::MDIndex<dd>& i) const {
            return idx(d, i[d]);
        }
        const auto& getDims() const { return dims; }
        int getDimOf(int d) const { return dims[d]; }
        auto getRange() const { return range; }
        auto getExtRange() const { return extRange; }
        auto getStart() const { return range.start; }
        auto getEnd() const { return range.end; }
        auto getStartOf(int d) const { return range.start[d]; }
        auto getEndOf(int d) const { return range.end[d]; }
        const auto& getView() const { return *this; }
        bool operator==(const CartesianMesh& o) const { return dims == o.dims && _x == o._x; }
    };

    template <typename M>
    struct MeshBuilder;

    template <int N>
    struct MeshBuilder<CartesianMesh<Meta::int_<N>>> {
        using Mesh = CartesianMesh<Meta::int_<N>>;
        static constexpr int dim = N;
        std::array<int, N> dims {}, start {};
        int padding_width = 5;// CartesianMesh.hpp:127
        std::array<MeshExtMode, N> ext_mode {};
        struct Axis {
            int kind = 0;// 0 unset, 1 uniform, 2 coordinates
            Real lo = 0, hi = 0;
            std::function<Real(int)> f;
        };
        std::array<Axis, N> axes;

        MeshBuilder() = default;
        template <typename... I>
        auto& newMesh(I... n) {// C-variadic in the reference (CartesianMesh.hpp:132-139)
            static_assert(sizeof...(I) == N, "newMesh: one size per dimension");
            dims = {static_cast<int>(n)...};
            return *this;
        }
        auto& newMesh(const Mesh& m) {
            dims = m.dims;
            start = m.range.start;
            for (int d = 0; d < N; ++d) {
                axes[d].kind = 2;
                auto xs = m._x[d];
                const int off = m.range.start[d] - m.extRange.start[d], s = m.range.start[d];
                axes[d].f = [xs, off, s](int i) { return xs[i - s + off]; };
            }
            return *this;
        }
        auto& setStart(const std::array<int, N>& s) {
            start = s;
            return *this;
        }
        auto& setPadWidth(int w) {
            padding_width = w;
            return *this;
        }
        auto& setExtMode(MeshExtMode m) {
            ext_mode.fill(m);
            return *this;
        }
        auto& setExtMode(int d, MeshExtMode m) {
            ext_mode[d] = m;
            return *this;
        }
        auto& setMeshOfDim(int k, Real lo, Real hi) {// CartesianMesh.hpp:168-171
            axes[k].kind = 1;
            axes[k].lo = lo;
            axes[k].hi = hi;
            return *this;
        }
        template <typename F>
        requires std::is_invocable_r_v<Real, F, int> auto& setMeshOfDim(int k, F&& f) {// :163-166
            axes[k].kind = 2;
            axes[k].f = std::forward<F>(f);
            return *this;
        }
        Mesh build() {
            Mesh m;
            opf_mesh_t h = internal::check_ptr(opf_mesh_create(N, dims.data(), start.data(), padding_width), "opf_mesh_create");
            m.handle = std::make_shared<internal::MeshHandle>(h);
            m.dims = dims;
            for (int d = 0; d < N; ++d) {
                if (ext_mode[d] != MeshExtMode::Undefined) internal::check_rc(opf_mesh_set_ext_mode(h, d, static_cast<int>(ext_mode[d])), "opf_mesh_set_ext_mode");
                if (axes[d].kind == 1) internal::check_rc(opf_mesh_set_uniform(h, d, axes[d].lo, axes[d].hi), "opf_mesh_set_uniform");
                else if (axes[d].kind == 2) {
                    std::vector<Real> xs(dims[d]);
                    for (int i = 0; i < dims[d]; ++i) xs[i] = axes[d].f(start[d] + i);
                    internal::check_rc(opf_mesh_set_coords(h, d, xs.data(), dims[d]), "opf_mesh_set_coords");
                } else {
                    OP_CRITICAL("MeshBuilder: setMeshOfDim({}, ...) missing", d);
                    OP_ABORT;
                }
            }
            opf_range r, e;
            internal::check_rc(opf_mesh_get_range(h, &r, &e), "opf_mesh_get_range");
            m.range = internal::from_c<N>(r);
            m.extRange = internal::from_c<N>(e);
            for (int d = 0; d < N; ++d) {
                const int n = m.extRange.end[d] - m.extRange.start[d];
                m._x[d].assign(n, 0.);
                m._dx[d].assign(n - 1, 0.);
                m._idx[d].assign(n - 1, 0.);
                if (opf_mesh_get_axis(h, d, m._x[d].data(), m._dx[d].data(), m._idx[d].data(), n) != n) {
                    OP_CRITICAL("opf_mesh_get_axis failed");
                    OP_ABORT;
                }
            }
            return m;
        }
    };
}// namespace OpFlow
