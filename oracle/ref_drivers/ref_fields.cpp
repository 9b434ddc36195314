// oracle/ref_drivers/ref_fields.cpp -- TEST INFRASTRUCTURE.
// Emits golden data from the UNMODIFIED reference for the index maps / boundary classification of the hot path:
//   * field ranges for every {loc} x {BC} x {ext} combination of ExprBuilder::calculateRanges (CartesianField.hpp:950-1029)
//   * ghost values after updatePadding() for Dirc/Neum/Symm/ASymm/Periodic on uniform and stretched meshes
//   * ranges / loc of prepared expressions (Op::prepare of every stencil operator)
//   * EvenSplitStrategy::getSplitMap for several rank counts (EvenSplitStrategy.hpp:57-192)
// Output: one JSON document on stdout (consumed by oracle/make_golden.py -> tests/golden/ref_fields.json).
#include "ref_common.hpp"
#include <sstream>
using namespace OpFlow;
using namespace refdrv;

template <std::size_t d>
static std::string rj(const DS::Range<d>& r) {
    std::ostringstream o;
    o << "[[";
    for (std::size_t i = 0; i < d; ++i) o << (i ? "," : "") << r.start[i];
    o << "],[";
    for (std::size_t i = 0; i < d; ++i) o << (i ? "," : "") << r.end[i];
    o << "]]";
    return o.str();
}
template <typename F>
static std::string ranges_json(const F& u) {
    std::ostringstream o;
    o << "{\"local\":" << rj(u.localRange) << ",\"assignable\":" << rj(u.assignableRange) << ",\"accessible\":" << rj(u.accessibleRange)
      << ",\"logical\":" << rj(u.logicalRange) << ",\"padding\":" << u.padding << "}";
    return o.str();
}
template <typename F>
static std::string values_json(const F& u) {
    std::ostringstream o;
    o.precision(17);
    auto r = u.getLocalReadableRange();
    o << "{\"range\":" << rj(r) << ",\"values\":[";
    bool first = true;
    rangeFor_s(r, [&](auto&& i) {
        o << (first ? "" : ",") << u.evalAt(i);
        first = false;
    });
    o << "]}";
    return o.str();
}
static const char* bcname(BCType t) {
    switch (t) {
        case BCType::Dirc: return "Dirc";
        case BCType::Neum: return "Neum";
        case BCType::Periodic: return "Periodic";
        case BCType::Symm: return "Symm";
        case BCType::ASymm: return "ASymm";
        default: return "Undefined";
    }
}
static double stretched(int i, int n) {
    double s = double(i) / (n - 1);
    return s + 0.15 * std::sin(2 * PI * s) / (2 * PI);
}

int main() {
    using Mesh1 = CartesianMesh<Meta::int_<1>>;
    using Field1 = CartesianField<Real, Mesh1>;
    using Mesh2 = CartesianMesh<Meta::int_<2>>;
    using Field2 = CartesianField<Real, Mesh2>;
    using Mesh3 = CartesianMesh<Meta::int_<3>>;
    using Field3 = CartesianField<Real, Mesh3>;
    std::ostringstream out;
    out.precision(17);
    out << "{\n";
    // ---------------------------------------------------------------- 1-D range classification table + ghost values
    out << "\"ranges1d\": [\n";
    {
        const int n = 11;
        auto mesh = MeshBuilder<Mesh1>().newMesh(n).setMeshOfDim(0, 0., 2.).build();
        auto smesh = MeshBuilder<Mesh1>().newMesh(n).setMeshOfDim(0, [&](int i) { return stretched(i, n); }).build();
        bool first = true;
        const BCType types[] = {BCType::Undefined, BCType::Dirc, BCType::Neum, BCType::Symm, BCType::ASymm, BCType::Periodic};
        for (int loc = 0; loc < 2; ++loc)
            for (auto ts : types)
                for (auto te : types)
                    for (int ext = 0; ext <= 2; ++ext)
                        for (int st = 0; st < 2; ++st) {
                            if ((ts == BCType::Periodic) != (te == BCType::Periodic)) continue;
                            auto b = ExprBuilder<Field1>().setMesh(st ? smesh : mesh).setLoc(loc ? LocOnMesh::Center : LocOnMesh::Corner).setExt(ext);
                            auto setbc = [&](DimPos pos, BCType t, double v) {
                                if (t == BCType::Undefined) return;
                                if (t == BCType::Dirc || t == BCType::Neum) b.setBC(0, pos, t, v);
                                else
                                    b.setBC(0, pos, t);
                            };
                            setbc(DimPos::start, ts, 1.25);
                            setbc(DimPos::end, te, -0.5);
                            auto u = b.build();
                            u.initBy([](auto&& x) { return 1.0 + 0.5 * x[0] + x[0] * x[0]; });
                            out << (first ? "" : ",\n") << "{\"n\":" << n << ",\"loc\":" << loc << ",\"bc\":[\"" << bcname(ts) << "\",\"" << bcname(te)
                                << "\"],\"bcv\":[1.25,-0.5],\"ext\":" << ext << ",\"stretched\":" << st << ",\"ranges\":" << ranges_json(u)
                                << ",\"field\":" << values_json(u) << "}";
                            first = false;
                        }
    }
    out << "\n],\n";
    // ---------------------------------------------------------------- 2-D ghost fill incl. corners (axis order)
    out << "\"ghost2d\": [\n";
    {
        const int nx = 9, ny = 7;
        bool first = true;
        for (int st = 0; st < 2; ++st)
            for (int locx = 0; locx < 2; ++locx)
                for (int locy = 0; locy < 2; ++locy)
                    for (int combo = 0; combo < 4; ++combo) {
                        auto mb = MeshBuilder<Mesh2>().newMesh(nx, ny);
                        if (st) mb.setMeshOfDim(0, [&](int i) { return 2 * stretched(i, nx); }).setMeshOfDim(1, [&](int i) { return stretched(i, ny); });
                        else
                            mb.setMeshOfDim(0, 0., 2.).setMeshOfDim(1, 0., 1.);
                        auto mesh = mb.build();
                        auto b = ExprBuilder<Field2>().setMesh(mesh).setLoc({locx ? LocOnMesh::Center : LocOnMesh::Corner, locy ? LocOnMesh::Center : LocOnMesh::Corner}).setExt(2);
                        std::string desc;
                        if (combo == 0) {
                            b.setBC(0, DimPos::start, BCType::Dirc, 1.).setBC(0, DimPos::end, BCType::Dirc, 2.).setBC(1, DimPos::start, BCType::Dirc, 3.).setBC(1, DimPos::end, BCType::Dirc, 4.);
                            desc = "[[\"Dirc\",1,\"Dirc\",2],[\"Dirc\",3,\"Dirc\",4]]";
                        } else if (combo == 1) {
                            b.setBC(0, DimPos::start, BCType::Neum, .5).setBC(0, DimPos::end, BCType::Dirc, 2.).setBC(1, DimPos::start, BCType::Dirc, 3.).setBC(1, DimPos::end, BCType::Neum, -1.);
                            desc = "[[\"Neum\",0.5,\"Dirc\",2],[\"Dirc\",3,\"Neum\",-1]]";
                        } else if (combo == 2) {
                            b.setBC(0, DimPos::start, BCType::Periodic).setBC(0, DimPos::end, BCType::Periodic).setBC(1, DimPos::start, BCType::Symm).setBC(1, DimPos::end, BCType::ASymm);
                            desc = "[[\"Periodic\",0,\"Periodic\",0],[\"Symm\",0,\"ASymm\",0]]";
                        } else {
                            b.setBC(0, DimPos::start, BCType::Periodic).setBC(0, DimPos::end, BCType::Periodic).setBC(1, DimPos::start, BCType::Periodic).setBC(1, DimPos::end, BCType::Periodic);
                            desc = "[[\"Periodic\",0,\"Periodic\",0],[\"Periodic\",0,\"Periodic\",0]]";
                        }
                        auto u = b.build();
                        u.initBy([](auto&& x) { return std::sin(1.3 * x[0]) + x[1] * x[1] + 0.25 * x[0] * x[1]; });
                        out << (first ? "" : ",\n") << "{\"dims\":[" << nx << "," << ny << "],\"loc\":[" << locx << "," << locy << "],\"bc\":" << desc
                            << ",\"ext\":2,\"stretched\":" << st << ",\"ranges\":" << ranges_json(u) << ",\"field\":" << values_json(u) << "}";
                        first = false;
                    }
    }
    out << "\n],\n";
    // ---------------------------------------------------------------- prepared-expression ranges / loc
    out << "\"prepare2d\": [\n";
    {
        const int nx = 12, ny = 10;
        auto mesh = MeshBuilder<Mesh2>().newMesh(nx, ny).setMeshOfDim(0, 0., 2.).setMeshOfDim(1, 0., 1.).build();
        bool first = true;
        for (int locx = 0; locx < 2; ++locx)
            for (int locy = 0; locy < 2; ++locy) {
                auto u = ExprBuilder<Field2>().setMesh(mesh).setLoc({locx ? LocOnMesh::Center : LocOnMesh::Corner, locy ? LocOnMesh::Center : LocOnMesh::Corner})
                                 .setBC(0, DimPos::start, BCType::Dirc, 1.).setBC(0, DimPos::end, BCType::Neum, 0.)
                                 .setBC(1, DimPos::start, BCType::Neum, 0.).setBC(1, DimPos::end, BCType::Dirc, 0.).setExt(3).build();
                auto emit = [&](const char* sig, auto&& e) {
                    e.prepare();
                    out << (first ? "" : ",\n") << "{\"loc\":[" << locx << "," << locy << "],\"sig\":\"" << sig << "\",\"acc\":" << rj(e.accessibleRange)
                        << ",\"local\":" << rj(e.localRange) << ",\"logical\":" << rj(e.logicalRange) << ",\"eloc\":[" << (e.loc[0] == LocOnMesh::Center) << ","
                        << (e.loc[1] == LocOnMesh::Center) << "]}";
                    first = false;
                };
                emit("D2C<0,F<0>>", d2x<D2SecondOrderCentered>(u));
                emit("D2C<1,F<0>>", d2y<D2SecondOrderCentered>(u));
                emit("D1C<0,F<0>>", dx<D1FirstOrderCentered>(u));
                emit("D1C<1,F<0>>", dy<D1FirstOrderCentered>(u));
                emit("D1Dn<0,F<0>>", dx<D1FirstOrderBiasedDownwind>(u));
                emit("D1Up<1,F<0>>", dy<D1FirstOrderBiasedUpwind>(u));
                emit("WenoDn<0,F<0>>", dx<D1WENO53Downwind>(u));
                emit("WenoUp<1,F<0>>", dy<D1WENO53Upwind>(u));
                if (locx) emit("IntpC2N<0,F<0>>", d1IntpCenterToCorner<0>(u));
                else
                    emit("IntpN2C<0,F<0>>", d1IntpCornerToCenter<0>(u));
                emit("Add<F<0>,Mul<S<0>,Add<D2C<0,F<1>>,D2C<1,F<2>>>>>", u + 0.1 * (d2x<D2SecondOrderCentered>(u) + d2y<D2SecondOrderCentered>(u)));
                emit("Sub<F<0>,Mul<S<0>,WenoDn<0,F<1>>>>", u - 0.1 * dx<D1WENO53Downwind>(u));
            }
    }
    out << "\n],\n";
    // ---------------------------------------------------------------- EvenSplitStrategy golden maps
    out << "\"split\": [\n";
    {
        bool first = true;
        auto emit2 = [&](int nx, int ny, int p) {
            auto range = DS::Range<2> {std::array<int, 2> {nx, ny}};
            auto plan = ParallelPlan {};
            plan.distributed_workers_count = p;
            setGlobalParallelPlan(plan);
            auto strategy = EvenSplitStrategy<Field2> {};
            auto map = strategy.getSplitMap(range, plan);
            out << (first ? "" : ",\n") << "{\"dim\":2,\"mesh\":[" << nx << "," << ny << "],\"ranks\":" << p << ",\"map\":[";
            for (size_t i = 0; i < map.size(); ++i) out << (i ? "," : "") << rj(map[i]);
            out << "]}";
            first = false;
        };
        auto emit3 = [&](int nx, int ny, int nz, int p) {
            auto range = DS::Range<3> {std::array<int, 3> {nx, ny, nz}};
            auto plan = ParallelPlan {};
            plan.distributed_workers_count = p;
            setGlobalParallelPlan(plan);
            auto strategy = EvenSplitStrategy<Field3> {};
            auto map = strategy.getSplitMap(range, plan);
            out << (first ? "" : ",\n") << "{\"dim\":3,\"mesh\":[" << nx << "," << ny << "," << nz << "],\"ranks\":" << p << ",\"map\":[";
            for (size_t i = 0; i < map.size(); ++i) out << (i ? "," : "") << rj(map[i]);
            out << "]}";
            first = false;
        };
        emit2(33, 33, 4), emit2(49, 33, 6), emit2(34, 34, 4), emit2(50, 34, 6), emit2(1025, 1025, 8), emit2(4097, 4097, 2), emit2(65, 129, 3), emit2(17, 1025, 8);
        emit3(33, 33, 33, 8), emit3(1025, 1025, 1025, 8), emit3(513, 513, 513, 4), emit3(65, 33, 17, 6), emit3(129, 129, 1025, 8), emit3(1025, 1025, 1025, 2);
        auto plan = ParallelPlan {};
        setGlobalParallelPlan(plan);
    }
    out << "\n]\n}\n";
    fputs(out.str().c_str(), stdout);
    return 0;
}
