// tests/frontend/fe_io.cpp -- binary stream writers / readers of the front-end: three time steps streamed through RawBinaryOStream (the
// reference's .cart layout, RawBinaryStream.hpp:98-227) and H5Stream with asynchronous device snapshots, read back through
// RawBinaryIStream / H5Stream(StreamIn) into another field: bit-identical to the field that was written.  Prints PASS.
#include <OpFlow>
using namespace OpFlow;
int main(int argc, char** argv) {
    EnvironmentGardian _(&argc, &argv);
    using Mesh = CartesianMesh<Meta::int_<2>>;
    using Field = CartesianField<Real, Mesh>;
    auto mesh = MeshBuilder<Mesh>().newMesh(33, 17).setMeshOfDim(0, 0., 1.).setMeshOfDim(1, 0., 1.).build();
    auto u = ExprBuilder<Field>().setName("u").setMesh(mesh).setBC(0, DimPos::start, BCType::Dirc, 1.).setBC(0, DimPos::end, BCType::Dirc, 1.)
                     .setBC(1, DimPos::start, BCType::Dirc, 1.).setBC(1, DimPos::end, BCType::Dirc, 1.).build();
    u.initBy([](auto&& x) { return std::sin(3 * x[0]) + x[1]; });
    auto v = u;
    v.name = "u";
    {
        Utils::RawBinaryOStream os("./out/u.cart");
        Utils::H5Stream hs("./out/sol.h5");
        for (int i = 0; i < 3; ++i) {
            u = u + 0.0001 * (d2x<D2SecondOrderCentered>(u) + d2y<D2SecondOrderCentered>(u));
            os << Utils::TimeStamp(i) << u;
            hs << Utils::TimeStamp(i) << u;
        }
    }
    v = 0;
    Utils::RawBinaryIStream is("./out/u.cart");
    is.setCounterTo(2);
    is >> v;
    double err = 0;
    rangeFor_s(u.localRange, [&](auto&& i) { err = std::max(err, std::abs(u[i] - v[i])); });
    v = 0;
    Utils::H5Stream hin("./out/sol.h5", Utils::StreamIn);
    hin.moveToTime(Utils::TimeStamp(2)) >> v;
    rangeFor_s(u.localRange, [&](auto&& i) { err = std::max(err, std::abs(u[i] - v[i])); });
    std::printf(err == 0 ? "PASS\n" : "FAIL %g\n", err);
    return err == 0 ? 0 : 1;
}
