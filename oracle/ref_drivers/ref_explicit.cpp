// oracle/ref_drivers/ref_explicit.cpp -- TEST INFRASTRUCTURE.
// Runs the reference's explicit path (Field = expr) for the BASELINE configs C1/C2/C3 and prints
// one JSON line with timing; optionally dumps the final field.  The expressions are the ones of
// examples/FTCS2D/FTCS-OMP.cpp:26 (C1), its d2z extension (C2) and examples/CONV1D/CONV1D.cpp:29-31
// with D1WENO53Downwind/Upwind or D1FirstOrderBiasedDownwind (C3).
//
//   ref_explicit --case ftcs2d|ftcs3d|ftcs2d_mpi|ftcs2d_fbc|weno_down|weno_up|upwind1 --n N --steps S --warmup W
//                --threads T --init zero|sin --dump path --ghosts 0|1
#include "ref_common.hpp"
using namespace OpFlow;
using namespace refdrv;

template <typename F>
static void finish(const char* name, int n, int steps, int threads, double sec, long long cells, const F& u,
                   const char* dump, bool ghosts) {
    printf("{\"case\": \"%s\", \"n\": %d, \"steps\": %d, \"threads\": %d, \"seconds\": %.6f, \"cells_per_step\": %lld, "
           "\"mlups\": %.3f}\n",
           name, n, steps, threads, sec, cells, steps > 0 ? cells * (double) steps / sec / 1e6 : 0.0);
    if (dump && *dump) dump_field(dump, u, ghosts);
}

int main(int argc, char** argv) {
    EnvironmentGardian _env(&argc, &argv);
    std::string cs = arg(argc, argv, "--case", "ftcs2d");
    int n = atoi(arg(argc, argv, "--n", "65"));
    int steps = atoi(arg(argc, argv, "--steps", "10"));
    int warm = atoi(arg(argc, argv, "--warmup", "0"));
    int nt = atoi(arg(argc, argv, "--threads", "1"));
    std::string init = arg(argc, argv, "--init", "zero");
    const char* dump = arg(argc, argv, "--dump", "");
    bool ghosts = atoi(arg(argc, argv, "--ghosts", "0"));
    set_threads(nt);

    if (cs == "ftcs2d") {
        using Mesh = CartesianMesh<Meta::int_<2>>;
        using Field = CartesianField<Real, Mesh>;
        auto mesh = MeshBuilder<Mesh>().newMesh(n, n).setMeshOfDim(0, 0., 1.).setMeshOfDim(1, 0., 1.).build();
        auto u = ExprBuilder<Field>().setName("u").setMesh(mesh)
                         .setBC(0, DimPos::start, BCType::Dirc, 1.).setBC(0, DimPos::end, BCType::Dirc, 1.)
                         .setBC(1, DimPos::start, BCType::Dirc, 1.).setBC(1, DimPos::end, BCType::Dirc, 1.)
                         .build();
        if (init == "sin") u.initBy([](auto&& x) { return std::sin(PI * x[0]) * std::sin(PI * x[1]); });
        else u = 0;
        const Real dt = 0.1 / Math::pow2(n - 1), alpha = 1.0;
        auto step = [&] { u = u + dt * alpha * (d2x<D2SecondOrderCentered>(u) + d2y<D2SecondOrderCentered>(u)); };
        for (int i = 0; i < warm; ++i) step();
        double t0 = now();
        for (int i = 0; i < steps; ++i) step();
        double t1 = now();
        finish("ftcs2d", n, steps, nt, t1 - t0, (long long) (n - 2) * (n - 2), u, dump, ghosts);
    } else if (cs == "ftcs3d") {
        using Mesh = CartesianMesh<Meta::int_<3>>;
        using Field = CartesianField<Real, Mesh>;
        auto mesh = MeshBuilder<Mesh>().newMesh(n, n, n).setMeshOfDim(0, 0., 1.).setMeshOfDim(1, 0., 1.)
                            .setMeshOfDim(2, 0., 1.).build();
        auto u = ExprBuilder<Field>().setName("u").setMesh(mesh)
                         .setBC(0, DimPos::start, BCType::Dirc, 1.).setBC(0, DimPos::end, BCType::Dirc, 1.)
                         .setBC(1, DimPos::start, BCType::Dirc, 1.).setBC(1, DimPos::end, BCType::Dirc, 1.)
                         .setBC(2, DimPos::start, BCType::Dirc, 1.).setBC(2, DimPos::end, BCType::Dirc, 1.)
                         .build();
        if (init == "sin")
            u.initBy([](auto&& x) { return std::sin(PI * x[0]) * std::sin(PI * x[1]) * std::sin(PI * x[2]); });
        else u = 0;
        const Real dt = 0.1 / Math::pow2(n - 1), alpha = 1.0;
        auto step = [&] {
            u = u + dt * alpha * (d2x<D2SecondOrderCentered>(u) + d2y<D2SecondOrderCentered>(u) + d2z<D2SecondOrderCentered>(u));
        };
        for (int i = 0; i < warm; ++i) step();
        double t0 = now();
        for (int i = 0; i < steps; ++i) step();
        double t1 = now();
        finish("ftcs3d", n, steps, nt, t1 - t0, (long long) (n - 2) * (n - 2) * (n - 2), u, dump, ghosts);
    } else if (cs == "ftcs2d_fbc") {
        // functor boundary conditions (setBC(d, pos, type, functor), CartesianField.hpp:870-892; FunctorDircBC DircBC.hpp:83-118):
        // Dirichlet values varying along the x faces, a Neumann flux varying along the upper y face, a constant on the lower one
        using Mesh = CartesianMesh<Meta::int_<2>>;
        using Field = CartesianField<Real, Mesh>;
        auto mesh = MeshBuilder<Mesh>().newMesh(n, n).setMeshOfDim(0, 0., 1.).setMeshOfDim(1, 0., 2.).build();
        const Real h = 2. / (n - 1);
        auto u = ExprBuilder<Field>().setName("u").setMesh(mesh)
                         .setBC(0, DimPos::start, BCType::Dirc, [=](auto&& i) { return 1. + 0.5 * std::sin(3. * h * i[1]); })
                         .setBC(0, DimPos::end, BCType::Dirc, [=](auto&& i) { return 0.25 * h * i[1]; })
                         .setBC(1, DimPos::start, BCType::Dirc, 0.75)
                         .setBC(1, DimPos::end, BCType::Neum, [=](auto&& i) { return std::cos(0.5 * h * i[0]); })
                         .setLoc(std::array {LocOnMesh::Center, LocOnMesh::Center}).setExt(1).build();
        if (init == "sin") u.initBy([](auto&& x) { return std::sin(PI * x[0]) * std::sin(PI * x[1]); });
        else u = 0;
        const Real dt = 0.1 / Math::pow2(n - 1), alpha = 1.0;
        auto step = [&] { u = u + dt * alpha * (d2x<D2SecondOrderCentered>(u) + d2y<D2SecondOrderCentered>(u)); };
        for (int i = 0; i < warm; ++i) step();
        double t0 = now();
        for (int i = 0; i < steps; ++i) step();
        double t1 = now();
        finish("ftcs2d_fbc", n, steps, nt, t1 - t0, (long long) (n - 1) * (n - 1), u, dump, ghosts);
    } else if (cs == "ftcs2d_mpi") {
        // the set-up of examples/FTCS2D/FTCS-MPI.cpp:12-40: cell-centred field, ext 1, padding 1, EvenSplitStrategy over the
        // distributed workers of the global plan (one worker here for the reference build; N GPUs for the B200 front-end)
        using Mesh = CartesianMesh<Meta::int_<2>>;
        using Field = CartesianField<Real, Mesh>;
        auto mesh = MeshBuilder<Mesh>().newMesh(n, n).setMeshOfDim(0, 0., 1.).setMeshOfDim(1, 0., 1.).build();
        auto info = makeParallelInfo();
        info.threadInfo.thread_count = nt;
        setGlobalParallelInfo(info);
        setGlobalParallelPlan(makeParallelPlan(getGlobalParallelInfo(), ParallelIdentifier::DistributeMem | ParallelIdentifier::SharedMem));
        std::shared_ptr<AbstractSplitStrategy<Field>> strategy = std::make_shared<EvenSplitStrategy<Field>>();
        auto u = ExprBuilder<Field>().setName("u").setMesh(mesh)
                         .setBC(0, DimPos::start, BCType::Dirc, 1.).setBC(0, DimPos::end, BCType::Dirc, 1.)
                         .setBC(1, DimPos::start, BCType::Dirc, 1.).setBC(1, DimPos::end, BCType::Dirc, 1.)
                         .setLoc(std::array {LocOnMesh::Center, LocOnMesh::Center}).setExt(1).setPadding(1)
                         .setSplitStrategy(strategy).build();
        if (init == "sin") u.initBy([](auto&& x) { return std::sin(PI * x[0]) * std::sin(PI * x[1]); });
        else u = 0;
        const Real dt = 0.1 / Math::pow2(n - 1), alpha = 1.0;
        auto step = [&] { u = u + dt * alpha * (d2x<D2SecondOrderCentered>(u) + d2y<D2SecondOrderCentered>(u)); };
        for (int i = 0; i < warm; ++i) step();
        double t0 = now();
        for (int i = 0; i < steps; ++i) step();
        double t1 = now();
        // every worker dumps its own block (file name gets the worker id when there are several)
        std::string path = dump;
        if (getWorkerCount() > 1 && !path.empty()) path += "." + std::to_string(getWorkerId());
        finish("ftcs2d_mpi", n, steps, nt, t1 - t0, (long long) (n - 1) * (n - 1), u, path.c_str(), ghosts);
    } else if (cs == "weno_down" || cs == "weno_up" || cs == "upwind1") {
        using Mesh = CartesianMesh<Meta::int_<1>>;
        using Field = CartesianField<Real, Mesh>;
        auto mesh = MeshBuilder<Mesh>().newMesh(n).setMeshOfDim(0, 0., 1.).build();
        auto u = ExprBuilder<Field>().setMesh(mesh).setName("u")
                         .setBC(0, DimPos::start, BCType::Dirc, 0.).setBC(0, DimPos::end, BCType::Dirc, 0.)
                         .setLoc(LocOnMesh::Corner).setExt(3).build();
        if (init == "sin") u.initBy([](auto&& x) { return std::sin(2 * PI * x[0]); });
        else u.initBy([](auto&& i) { return 0.2 <= i[0] && i[0] <= 0.4 ? 1.0 : 0.0; });
        const Real dt = 0.5 / (n - 1);
        const Real c = 1.0;
        auto step = [&] {
            if (cs == "weno_down") u = u - dt * c * dx<D1WENO53Downwind>(u);
            else if (cs == "weno_up") u = u + dt * c * dx<D1WENO53Upwind>(u);
            else u = u - dt * c * dx<D1FirstOrderBiasedDownwind>(u);
        };
        for (int i = 0; i < warm; ++i) step();
        double t0 = now();
        for (int i = 0; i < steps; ++i) step();
        double t1 = now();
        finish(cs.c_str(), n, steps, nt, t1 - t0, (long long) (n - 2), u, dump, ghosts);
    } else {
        fprintf(stderr, "unknown case %s\n", cs.c_str());
        return 2;
    }
    return 0;
}
