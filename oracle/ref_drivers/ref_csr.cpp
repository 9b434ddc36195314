// oracle/ref_drivers/ref_csr.cpp -- TEST INFRASTRUCTURE.
// Stencil coefficients of the UNMODIFIED reference on meshes whose spacing is NOT 1 (the reference's own
// CSRMatrixGeneratorTest.cpp:57-158 uses dx = 1, where every reciprocal is exact): the equation `1.0 == d2x(e) + d2y(e)`
// (CSRMatrixGeneratorTest.cpp:48-50) assembled by CSRMatrixGenerator::generate on 6 x 5 cells, uniform non-unit spacing and a
// stretched mesh, Dirichlet / Neumann (pinned) / periodic (pinned), Center and Corner location.  Values are printed as hex
// doubles so the JSON fixture is bit-exact (tests/golden/ref_csr.json).
#include "ref_common.hpp"
#include <cinttypes>
#include <sstream>
using namespace OpFlow;
using namespace refdrv;

static std::string hexd(double v) {
    char buf[40];
    snprintf(buf, sizeof buf, "\"%a\"", v);
    return buf;
}

int main() {
    using Mesh = CartesianMesh<Meta::int_<2>>;
    using Field = CartesianField<Real, Mesh>;
    set_threads(1);
    const int nx = 7, ny = 6;
    std::ostringstream out;
    out << "{\"cases\": [\n";
    bool first = true;
    for (int stretched = 0; stretched < 2; ++stretched) {
        std::vector<double> xs(nx), ys(ny);
        for (int i = 0; i < nx; ++i) {
            const double s = (double) i / (nx - 1);
            xs[i] = stretched ? 0.7 * (s + 0.15 * std::sin(2 * PI * s) / (2 * PI)) : 0.;
        }
        for (int i = 0; i < ny; ++i) {
            const double s = (double) i / (ny - 1);
            ys[i] = stretched ? 1.3 * (s + 0.15 * std::sin(2 * PI * s) / (2 * PI)) : 0.;
        }
        auto mb = MeshBuilder<Mesh>().newMesh(nx, ny);
        Mesh m = stretched ? mb.setMeshOfDim(0, [&](int i) { return xs[i]; }).setMeshOfDim(1, [&](int i) { return ys[i]; }).build()
                           : mb.setMeshOfDim(0, 0., 0.7).setMeshOfDim(1, 0., 1.3).build();
        for (int loc = 0; loc < 2; ++loc)
            for (int bc = 0; bc < 3; ++bc) {
                if (bc == 2 && stretched) continue;// periodic needs a periodic mesh extension
                auto b = ExprBuilder<Field>().setMesh(m).setName("p").setExt(1);
                b.setLoc(loc ? std::array {LocOnMesh::Center, LocOnMesh::Center} : std::array {LocOnMesh::Corner, LocOnMesh::Corner});
                for (int d = 0; d < 2; ++d) {
                    if (bc == 0) b.setBC(d, DimPos::start, BCType::Dirc, 0.25 * (d + 1)).setBC(d, DimPos::end, BCType::Dirc, -0.5);
                    else if (bc == 1)
                        b.setBC(d, DimPos::start, BCType::Neum, 0.125).setBC(d, DimPos::end, BCType::Neum, 0.);
                    else
                        b.setBC(d, DimPos::start, BCType::Periodic).setBC(d, DimPos::end, BCType::Periodic);
                }
                Field p = b.build();
                auto eqn_f = [&](auto&& e) { return 1.0 == d2x<D2SecondOrderCentered>(e) + d2y<D2SecondOrderCentered>(e); };
                auto eqn = makeEqnHolder(std::forward_as_tuple(eqn_f), std::forward_as_tuple(p));
                auto st = makeStencilHolder(eqn);
                auto mat = CSRMatrixGenerator::generate<0>(st, DS::ColoredMDRangeMapper<2> {p.assignableRange}, bc != 0);
                out << (first ? "" : ",\n") << "{\"stretched\":" << stretched << ",\"loc\":" << loc << ",\"bc\":" << bc << ",\"pinned_last\":" << (bc != 0)
                    << ",\"dims\":[" << nx << "," << ny << "],\"range\":[[" << p.assignableRange.start[0] << "," << p.assignableRange.start[1] << "],["
                    << p.assignableRange.end[0] << "," << p.assignableRange.end[1] << "]],\"ptr\":[";
                first = false;
                for (size_t i = 0; i < mat.row.size(); ++i) out << (i ? "," : "") << mat.row[i];
                out << "],\"col\":[";
                for (size_t i = 0; i < mat.col.size(); ++i) out << (i ? "," : "") << mat.col[i];
                out << "],\"val\":[";
                for (size_t i = 0; i < mat.val.size(); ++i) out << (i ? "," : "") << hexd(mat.val[i]);
                out << "],\"rhs\":[";
                for (size_t i = 0; i < mat.rhs.size(); ++i) out << (i ? "," : "") << hexd(mat.rhs[i]);
                out << "]}";
            }
    }
    out << "\n]}\n";
    fputs(out.str().c_str(), stdout);
    return 0;
}
