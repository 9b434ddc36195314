// oracle/ref_drivers/ref_implicit.cpp -- TEST INFRASTRUCTURE.
// Implicit solves by the UNMODIFIED reference (HYPREEqnSolveHandler + vendored HYPRE 2.33.0), shaped after
// test/Core/Equation/{Dirc,Neum,Periodic}EqnTest.cpp: manufactured solution p_true, b = L_h(p_true) built with the explicit
// path, Solve(L_h(e) == b) driven to a tight tolerance.  Emits JSON with b, the solved field and the solver state so the
// B200 engine's matrix-free solve can be compared on identical data (tests/golden/ref_implicit.json).
#include "ref_common.hpp"
#include <sstream>
using namespace OpFlow;
using namespace refdrv;

template <typename F>
static std::string vals(const F& u, const typename F::RangeType& r) {
    std::ostringstream o;
    o.precision(17);
    o << "[";
    bool first = true;
    rangeFor_s(r, [&](auto&& i) {
        o << (first ? "" : ",") << u.evalAt(i);
        first = false;
    });
    o << "]";
    return o.str();
}
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

int main() {
    using Mesh2 = CartesianMesh<Meta::int_<2>>;
    using Field2 = CartesianField<Real, Mesh2>;
    using Mesh3 = CartesianMesh<Meta::int_<3>>;
    using Field3 = CartesianField<Real, Mesh3>;
    set_threads(4);
    std::ostringstream out;
    out.precision(17);
    out << "{\n\"cases\": [\n";
    bool first = true;
    auto emit = [&](const char* name, auto& p, auto& b, const EqnSolveState& st, const char* extra) {
        out << (first ? "" : ",\n") << "{\"name\":\"" << name << "\"," << extra << "\"range\":" << rj(p.assignableRange) << ",\"b\":" << vals(b, p.assignableRange)
            << ",\"p\":" << vals(p, p.assignableRange) << ",\"niter\":" << st.niter << ",\"relerr\":" << st.relerr << "}";
        first = false;
    };
    // GMRES + PFMG like LidDriven2D.cpp:45-49 (HYPRE's PCG refuses the negative-definite d2x+d2y operator)
    StructSolverParams<StructSolverType::GMRES> pcg;
    pcg.tol = 1e-13;
    pcg.maxIter = 200;
    StructSolverParams<StructSolverType::PFMG> pfmg;
    // ---- A: Dirichlet, cell-centred, 65^2 (DircEqnTest.cpp:30-62 geometry), constant-coefficient Laplacian
    {
        auto m = MeshBuilder<Mesh2>().newMesh(65, 65).setMeshOfDim(0, 0., 1.).setMeshOfDim(1, 0., 1.).build();
        auto p = ExprBuilder<Field2>().setMesh(m).setName("p").setBC(0, DimPos::start, BCType::Dirc, 0.).setBC(0, DimPos::end, BCType::Dirc, 0.)
                         .setBC(1, DimPos::start, BCType::Dirc, 0.).setBC(1, DimPos::end, BCType::Dirc, 0.).setExt(1)
                         .setLoc({LocOnMesh::Center, LocOnMesh::Center}).build();
        auto p_true = p;
        auto b = p;
        p_true.initBy([&](auto&& x) { return x[0] * (1. - x[0]) * x[1] * (1. - x[1]); });
        b = d2x<D2SecondOrderCentered>(p_true) + d2y<D2SecondOrderCentered>(p_true);
        p = 0.;
        auto st = Solve([&](auto&& e) { return d2x<D2SecondOrderCentered>(e) + d2y<D2SecondOrderCentered>(e) == b; }, p, pcg, pfmg);
        emit("dirc_center_65", p, b, st, "\"n\":[65,65],\"lo\":[0,0],\"hi\":[1,1],\"loc\":[1,1],\"bc\":\"Dirc\",\"bcv\":0,\"pin\":0,\"ext\":1,");
    }
    // ---- B: inhomogeneous Dirichlet, node-centred (Corner), 33x49, stretched box
    {
        auto m = MeshBuilder<Mesh2>().newMesh(33, 49).setMeshOfDim(0, 0., 1.).setMeshOfDim(1, 0., 2.).build();
        auto p = ExprBuilder<Field2>().setMesh(m).setName("p").setBC(0, DimPos::start, BCType::Dirc, 1.5).setBC(0, DimPos::end, BCType::Dirc, 1.5)
                         .setBC(1, DimPos::start, BCType::Dirc, 1.5).setBC(1, DimPos::end, BCType::Dirc, 1.5).setExt(1).build();
        auto p_true = p;
        auto b = p;
        p_true.initBy([&](auto&& x) { return 1.5 + std::sin(PI * x[0]) * std::sin(PI * x[1] / 2.); });
        b = d2x<D2SecondOrderCentered>(p_true) + d2y<D2SecondOrderCentered>(p_true);
        p = 0.;
        auto st = Solve([&](auto&& e) { return d2x<D2SecondOrderCentered>(e) + d2y<D2SecondOrderCentered>(e) == b; }, p, pcg, pfmg);
        emit("dirc_corner_33x49", p, b, st, "\"n\":[33,49],\"lo\":[0,0],\"hi\":[1,2],\"loc\":[0,0],\"bc\":\"Dirc\",\"bcv\":1.5,\"pin\":0,\"ext\":1,");
    }
    // ---- C: Neumann + pinned value, cell-centred, 33^2 on [0,2pi]^2 (NeumEqnTest.cpp:30-62), the LidDriven pressure set-up
    {
        auto m = MeshBuilder<Mesh2>().newMesh(33, 33).setMeshOfDim(0, 0., 2 * PI).setMeshOfDim(1, 0., 2 * PI).build();
        auto p = ExprBuilder<Field2>().setMesh(m).setName("p").setBC(0, DimPos::start, BCType::Neum, 0.).setBC(0, DimPos::end, BCType::Neum, 0.)
                         .setBC(1, DimPos::start, BCType::Neum, 0.).setBC(1, DimPos::end, BCType::Neum, 0.).setExt(1)
                         .setLoc({LocOnMesh::Center, LocOnMesh::Center}).build();
        auto p_true = p;
        auto b = p;
        p_true.initBy([&](auto&& x) { return std::cos(x[0]) * std::cos(x[1]); });
        b = d2x<D2SecondOrderCentered>(p_true) + d2y<D2SecondOrderCentered>(p_true);
        p = 0.;
        auto prm = pcg;
        prm.pinValue = true;
        auto st = Solve([&](auto&& e) { return d2x<D2SecondOrderCentered>(e) + d2y<D2SecondOrderCentered>(e) == b; }, p, prm, pfmg);
        emit("neum_center_33_pin", p, b, st, "\"n\":[33,33],\"lo\":[0,0],\"hi\":[6.283185307179586,6.283185307179586],\"loc\":[1,1],\"bc\":\"Neum\",\"bcv\":0,\"pin\":1,\"ext\":1,");
    }
    // ---- D: periodic + pinned value, cell-centred, 33^2 (PeriodicEqnTest.cpp:30-62)
    {
        auto m = MeshBuilder<Mesh2>().newMesh(33, 33).setMeshOfDim(0, 0., 2 * PI).setMeshOfDim(1, 0., 2 * PI).build();
        auto p = ExprBuilder<Field2>().setMesh(m).setName("p").setBC(0, DimPos::start, BCType::Periodic).setBC(0, DimPos::end, BCType::Periodic)
                         .setBC(1, DimPos::start, BCType::Periodic).setBC(1, DimPos::end, BCType::Periodic).setExt(1)
                         .setLoc({LocOnMesh::Center, LocOnMesh::Center}).build();
        auto p_true = p;
        auto b = p;
        p_true.initBy([&](auto&& x) { return std::sin(x[0]) * std::cos(x[1]); });
        b = d2x<D2SecondOrderCentered>(p_true) + d2y<D2SecondOrderCentered>(p_true);
        p = 0.;
        auto prm = pcg;
        prm.pinValue = true;
        auto st = Solve([&](auto&& e) { return d2x<D2SecondOrderCentered>(e) + d2y<D2SecondOrderCentered>(e) == b; }, p, prm, pfmg);
        emit("periodic_center_33_pin", p, b, st, "\"n\":[33,33],\"lo\":[0,0],\"hi\":[6.283185307179586,6.283185307179586],\"loc\":[1,1],\"bc\":\"Periodic\",\"bcv\":0,\"pin\":1,\"ext\":1,");
    }
    // ---- E: 3-D Dirichlet, node-centred, 17^3
    {
        auto m = MeshBuilder<Mesh3>().newMesh(17, 17, 17).setMeshOfDim(0, 0., 1.).setMeshOfDim(1, 0., 1.).setMeshOfDim(2, 0., 1.).build();
        auto bld = ExprBuilder<Field3>().setMesh(m).setName("p").setExt(1);
        for (int d = 0; d < 3; ++d) bld.setBC(d, DimPos::start, BCType::Dirc, 0.).setBC(d, DimPos::end, BCType::Dirc, 0.);
        auto p = bld.build();
        auto p_true = p;
        auto b = p;
        p_true.initBy([&](auto&& x) { return std::sin(PI * x[0]) * std::sin(PI * x[1]) * std::sin(PI * x[2]); });
        b = d2x<D2SecondOrderCentered>(p_true) + d2y<D2SecondOrderCentered>(p_true) + d2z<D2SecondOrderCentered>(p_true);
        p = 0.;
        auto st = Solve([&](auto&& e) { return d2x<D2SecondOrderCentered>(e) + d2y<D2SecondOrderCentered>(e) + d2z<D2SecondOrderCentered>(e) == b; }, p, pcg, pfmg);
        emit("dirc_corner_17cubed", p, b, st, "\"n\":[17,17,17],\"lo\":[0,0,0],\"hi\":[1,1,1],\"loc\":[0,0,0],\"bc\":\"Dirc\",\"bcv\":0,\"pin\":0,\"ext\":1,");
    }
    out << "\n]}\n";
    fputs(out.str().c_str(), stdout);
    return 0;
}
