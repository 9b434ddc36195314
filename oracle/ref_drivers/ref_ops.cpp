// oracle/ref_drivers/ref_ops.cpp -- TEST INFRASTRUCTURE.
// Values of the flux-limiter interpolators (src/Core/Operator/Interpolator/D1FluxLimiter.hpp, all ten schemes of
// D1FluxLimiterBasedIntpOp.hpp:22-61, both directions, along x and y) and of Convolution (Convolution.hpp) computed by the UNMODIFIED
// reference on a stretched 2-D mesh, printed as hex doubles with the prepared ranges (tests/golden/ref_ops.json).
// The advecting field changes sign inside the domain so that both the upwind and the downwind branch are exercised.
#include "ref_common.hpp"
#include <sstream>
using namespace OpFlow;
using namespace refdrv;
using Mesh = CartesianMesh<Meta::int_<2>>;
using Field = CartesianField<Real, Mesh>;

static std::ostringstream out;
static bool first = true;
static const Mesh* g_mesh = nullptr;

template <typename T>
static void emit(const char* name, int axis, const char* dir, T&& t) {
    t.prepare();
#ifdef OPFLOW_B200
    // the B200 front-end evaluates expressions on the device: assign into a field of the result's location, read that back
    auto dst = ExprBuilder<Field>().setMesh(*g_mesh).setLoc({t.loc[0], t.loc[1]}).setExt(2).build();
    dst = 0.;
    dst = t;
    auto value = [&](auto&& i) { return (double) dst.evalAt(i); };
#else
    auto value = [&](auto&& i) { return (double) t.evalAt(i); };
#endif
    out << (first ? "" : ",\n") << "{\"node\":\"" << name << "\",\"axis\":" << axis << ",\"dir\":\"" << dir << "\",\"acc\":[[" << t.accessibleRange.start[0] << ","
        << t.accessibleRange.start[1] << "],[" << t.accessibleRange.end[0] << "," << t.accessibleRange.end[1] << "]],\"local\":[[" << t.localRange.start[0] << ","
        << t.localRange.start[1] << "],[" << t.localRange.end[0] << "," << t.localRange.end[1] << "]],\"logical\":[[" << t.logicalRange.start[0] << ","
        << t.logicalRange.start[1] << "],[" << t.logicalRange.end[0] << "," << t.logicalRange.end[1] << "]],\"loc\":[" << (int) t.loc[0] << "," << (int) t.loc[1]
        << "],\"val\":[";
    first = false;
    bool f2 = true;
    rangeFor_s(t.accessibleRange, [&](auto&& i) {
        char buf[40];
        snprintf(buf, sizeof buf, "\"%a\"", value(i));
        out << (f2 ? "" : ",") << buf;
        f2 = false;
    });
    out << "]}";
}

#define ALL_SCHEMES(X) X(Central, D1Central) X(Quick, D1QUICK) X(Cui, D1CUI) X(Fromm, D1Fromm) X(Lui, D1LinearUpwind) X(Minmod, D1Minmod) \
    X(Superbee, D1Superbee) X(Muscl, D1MUSCL) X(Harmonic, D1Harmonic) X(Albada, D1Albada)

int main() {
    set_threads(1);
    const int nx = 14, ny = 11;
    auto sx = [&](int i) { double s = (double) i / (nx - 1); return 2.0 * (s + 0.15 * std::sin(2 * PI * s) / (2 * PI)); };
    auto sy = [&](int i) { double s = (double) i / (ny - 1); return 1.0 * (s + 0.15 * std::sin(2 * PI * s) / (2 * PI)); };
    auto m = MeshBuilder<Mesh>().newMesh(nx, ny).setMeshOfDim(0, sx).setMeshOfDim(1, sy).build();
    g_mesh = &m;
    auto mk = [&](LocOnMesh l0, LocOnMesh l1) {
        return ExprBuilder<Field>().setMesh(m).setBC(0, DimPos::start, BCType::Neum, 0.).setBC(0, DimPos::end, BCType::Neum, 0.)
                .setBC(1, DimPos::start, BCType::Neum, 0.).setBC(1, DimPos::end, BCType::Neum, 0.).setExt(2).setLoc({l0, l1}).build();
    };
    auto fe = [](auto&& x) { return std::sin(2.3 * x[0]) * std::cos(1.7 * x[1]) + 0.3 * x[0] * x[1]; };
    auto fu = [](auto&& x) { return std::cos(3.1 * x[0] + 0.4) * std::sin(2.2 * x[1] + 0.3); };
    out << "{\"nx\":" << nx << ",\"ny\":" << ny << ",\"cases\":[\n";
    {// along x
        auto e_c = mk(LocOnMesh::Center, LocOnMesh::Center), u_n = mk(LocOnMesh::Corner, LocOnMesh::Center);
        auto e_n = mk(LocOnMesh::Corner, LocOnMesh::Center), u_c = mk(LocOnMesh::Center, LocOnMesh::Center);
        e_c.initBy(fe), u_n.initBy(fu), e_n.initBy(fe), u_c.initBy(fu);
#define X(Name, Op)                                                                                                     \
    emit(#Name, 0, "C2N", d1IntpCenterToCorner<0, Op>(u_n, e_c));                                                       \
    emit(#Name, 0, "N2C", d1IntpCornerToCenter<0, Op>(u_c, e_n));
        ALL_SCHEMES(X)
#undef X
    }
    {// along y
        auto e_c = mk(LocOnMesh::Center, LocOnMesh::Center), u_n = mk(LocOnMesh::Center, LocOnMesh::Corner);
        auto e_n = mk(LocOnMesh::Center, LocOnMesh::Corner), u_c = mk(LocOnMesh::Center, LocOnMesh::Center);
        e_c.initBy(fe), u_n.initBy(fu), e_n.initBy(fe), u_c.initBy(fu);
        emit("Quick", 1, "C2N", d1IntpCenterToCorner<1, D1QUICK>(u_n, e_c));
        emit("Minmod", 1, "N2C", d1IntpCornerToCenter<1, D1Minmod>(u_c, e_n));
    }
    {// convolutions: 3 x 3 with the weights of UniLS.cpp:107-108, and an unequal 5 x 3 kernel with distinct entries
        auto e_c = mk(LocOnMesh::Center, LocOnMesh::Center);
        e_c.initBy(fe);
        constexpr auto _c = 16. / 24., _o = 1. / 24.;
        constexpr DS::FixedSizeTensor<double, 3, 3> k33 {_o, _o, _o, _o, _c, _o, _o, _o, _o};
        emit("Conv33", -1, "", conv(e_c, k33));
        DS::FixedSizeTensor<double, 5, 3> k53;
        for (int j = 0; j < 3; ++j)
            for (int i = 0; i < 5; ++i) k53[DS::MDIndex<2> {i, j}] = 0.1 * (i + 1) - 0.07 * (j + 1) * (i - 2);
        emit("Conv53", -1, "", conv(e_c, k53));
    }
    out << "\n]}\n";
    fputs(out.str().c_str(), stdout);
    return 0;
}
