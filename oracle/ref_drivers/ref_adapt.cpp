// oracle/ref_drivers/ref_adapt.cpp -- TEST INFRASTRUCTURE.
// The volume-constraint step of the level-set re-initialisation in examples/LevelSet/UniLS.cpp:104-118: an element-wise functor
// (UniOpAdaptor of Math::smoothDelta), pow, and 3 x 3 convolutions (the integral operator) composed in one assignment
//   p = p3 + lambda * (k + 1) * delta(p0),   lambda = -int(delta(p0) * (p3 - p0) / (k + 1)) / int(delta(p0)^2 + 1e-14)
// on a 65 x 65 cell-centred field.  Built against the unmodified reference (the functor is the example's constexpr lambda inside a
// NamedFunctor) and against the B200 front-end (the functor is a named type: nvcc cannot carry a lambda or a class-type template
// argument into a kernel, see opflow/field.hpp UniOpAdaptorT).  Dumps p (tests/golden/adapt_n65.opfd).
//   ref_adapt --n N --dump path
#include "ref_common.hpp"
using namespace OpFlow;
using namespace refdrv;

constexpr int N = 65;
constexpr double H = 1. / (N - 1);
#ifdef OPFLOW_B200
struct SmoothDelta {
    static constexpr const char* name = "smoothDelta";
    OPF_HD double operator()(double d) const { return Math::smoothDelta(H, d); }
};
#endif

int main(int argc, char** argv) {
    EnvironmentGardian _env(&argc, &argv);
    const char* dump = arg(argc, argv, "--dump", "");
    set_threads(1);
    using Mesh = CartesianMesh<Meta::int_<2>>;
    using Field = CartesianField<Real, Mesh>;
    auto mesh = MeshBuilder<Mesh>().newMesh(N, N).setMeshOfDim(0, 0., 1.).setMeshOfDim(1, 0., 1.).build();
    auto p0 = ExprBuilder<Field>().setName("p0").setMesh(mesh).setLoc({LocOnMesh::Center, LocOnMesh::Center}).setExt(1)
                      .setBC(0, DimPos::start, BCType::Neum, 0.).setBC(0, DimPos::end, BCType::Neum, 0.)
                      .setBC(1, DimPos::start, BCType::Neum, 0.).setBC(1, DimPos::end, BCType::Neum, 0.).build();
    auto p3 = p0, p = p0;
    p3.name = "p3";
    p.name = "p";
    p0.initBy([](auto&& x) { return std::sqrt((x[0] - .5) * (x[0] - .5) + (x[1] - .5) * (x[1] - .5)) - 0.25; });
    p3.initBy([](auto&& x) { return std::sqrt((x[0] - .5) * (x[0] - .5) + (x[1] - .5) * (x[1] - .5)) - 0.25 + 0.01 * std::sin(7 * x[0]) * std::cos(5 * x[1]); });
    p = 0;
    constexpr auto h = H;
    constexpr auto _c = 16. / 24., _o = 1. / 24.;
    constexpr DS::FixedSizeTensor<double, 3, 3> conv_ker {_o, _o, _o, _o, _c, _o, _o, _o, _o};
#ifdef OPFLOW_B200
    constexpr auto delta_op = [=](auto&& e) { return makeExpression<UniOpAdaptorT<SmoothDelta>>(OP_PERFECT_FOWD(e)); };
#else
    constexpr auto func = [=](Real d) { return Math::smoothDelta(h, d); };
    constexpr auto functor = Utils::NamedFunctor<func, Utils::makeCXprString("smoothDelta")>();
    constexpr auto delta_op = [=](auto&& e) { return makeExpression<UniOpAdaptor<functor>>(OP_PERFECT_FOWD(e)); };
#endif
    constexpr auto int_op = [=](auto&& e) { return h * h * conv(OP_PERFECT_FOWD(e), conv_ker); };
    const int k = 1;
    auto lambda = -int_op(delta_op(p0) * (p3 - p0) / (k + 1)) / int_op(pow(delta_op(p0), 2) + 1e-14);
    p = p3 + lambda * (k + 1) * delta_op(p0);
    double s = 0;
    rangeFor_s(p.assignableRange, [&](auto&& i) { s += p[i]; });
    printf("{\"case\": \"adapt\", \"n\": %d, \"sum\": %.17g}\n", N, s);
    if (dump && *dump) dump_field(dump, p, false);
    return 0;
}
