// oracle/ref_drivers/ref_tg3d.cpp -- TEST INFRASTRUCTURE and the source of BASELINE config C5.
// 3-D Taylor-Green vortex in a periodic box [0, 2 pi]^3 on a MAC-staggered mesh: the operator set and the time step of
// examples/LidDriven/LidDriven3D.cpp:59-160 (semi-implicit convection / diffusion momentum solves for du, dv, dw, explicit
// cross-term corrections, pressure Poisson solve with pinValue + staticMat, projection) with the periodic set-up of
// examples/TaylorGreen/TGMPI.cpp:35-49 (setExt(2), setPadding(2), split strategy) and the classical 3-D initial condition
// u = sin x cos y cos z, v = -cos x sin y cos z, w = 0, p = (cos 2x + cos 2y)(cos 2z + 2) / 16.
// The reference ships only a 2-D Taylor-Green (TG.cpp:18); SURVEY 8d composes the 3-D case exactly this way.
// Compiled twice from this one file: against the unmodified reference (oracle/build_ref.sh -> oracle/_ref/bin/ref_tg3d, HYPRE
// GMRES / PCG + PFMG) and against the B200 front-end (tests/frontend -> fe_tg3d, one process per GPU, z-slabs).
//   ref_tg3d --n N [--nz NZ] --steps S --threads T --tol 1e-12 --dump prefix
#include "ref_common.hpp"
using namespace OpFlow;
using namespace refdrv;

static void device_sync() {
#ifdef OPFLOW_B200
    opf_synchronize();
#endif
}

int main(int argc, char** argv) {
    EnvironmentGardian _env(&argc, &argv);
    using Mesh = CartesianMesh<Meta::int_<3>>;
    using Field = CartesianField<Real, Mesh>;
    const int n = atoi(arg(argc, argv, "--n", "33"));
    const int nz = atoi(arg(argc, argv, "--nz", "0")) > 0 ? atoi(arg(argc, argv, "--nz", "0")) : n;// nodes along z (weak scaling stretches z)
    const int steps = atoi(arg(argc, argv, "--steps", "2"));
    const int nt = atoi(arg(argc, argv, "--threads", "1"));
    const double tol = atof(arg(argc, argv, "--tol", "1e-12"));
    const std::string dump = arg(argc, argv, "--dump", "");
    auto info = makeParallelInfo();
    info.threadInfo.thread_count = nt;
    setGlobalParallelInfo(info);
    setGlobalParallelPlan(makeParallelPlan(getGlobalParallelInfo(), ParallelIdentifier::DistributeMem | ParallelIdentifier::SharedMem));
#ifdef OPFLOW_B200
    std::shared_ptr<AbstractSplitStrategy<Field>> strategy = std::make_shared<SlabSplitStrategy<Field>>();
#else
    std::shared_ptr<AbstractSplitStrategy<Field>> strategy = std::make_shared<EvenSplitStrategy<Field>>();
#endif
    const Real dt = 1e-3, nu = 1.0e-2;
    const Real lz = 2 * PI * (nz - 1) / (n - 1);// same spacing on every axis
    auto mesh = MeshBuilder<Mesh>().newMesh(n, n, nz).setMeshOfDim(0, 0., 2 * PI).setMeshOfDim(1, 0., 2 * PI).setMeshOfDim(2, 0., lz).build();
    auto builder = ExprBuilder<Field>().setMesh(mesh)
                           .setBC(0, DimPos::start, BCType::Periodic).setBC(0, DimPos::end, BCType::Periodic)
                           .setBC(1, DimPos::start, BCType::Periodic).setBC(1, DimPos::end, BCType::Periodic)
                           .setBC(2, DimPos::start, BCType::Periodic).setBC(2, DimPos::end, BCType::Periodic)
                           .setExt(2).setPadding(2).setSplitStrategy(strategy);
    auto u = builder.setName("u").setLoc({LocOnMesh::Corner, LocOnMesh::Center, LocOnMesh::Center}).build();
    auto du = u;
    du.name = "du";
    auto v = builder.setName("v").setLoc({LocOnMesh::Center, LocOnMesh::Corner, LocOnMesh::Center}).build();
    auto dv = v;
    dv.name = "dv";
    auto w = builder.setName("w").setLoc({LocOnMesh::Center, LocOnMesh::Center, LocOnMesh::Corner}).build();
    auto dw = w;
    dw.name = "dw";
    auto p = builder.setName("p").setLoc({LocOnMesh::Center, LocOnMesh::Center, LocOnMesh::Center}).build();
    auto dp = p;
    dp.name = "dp";
    u = 0; du = 0; v = 0; dv = 0; w = 0; dw = 0; p = 0; dp = 0;
    const Real kz = 2 * PI / lz;// one period along z whatever the box length
    u.initBy([&](auto&& x) { return std::sin(x[0]) * std::cos(x[1]) * std::cos(kz * x[2]); });
    v.initBy([&](auto&& x) { return -std::cos(x[0]) * std::sin(x[1]) * std::cos(kz * x[2]); });
    p.initBy([&](auto&& x) { return (std::cos(2 * x[0]) + std::cos(2 * x[1])) * (std::cos(2 * kz * x[2]) + 2.) / 16.; });

    // composite operators (LidDriven3D.cpp:59-88)
    auto conv_xx = [&](auto&& _1, auto&& _2) { return dx<D1FirstOrderCentered>(d1IntpCornerToCenter<0>(_1) * d1IntpCornerToCenter<0>(_2)); };
    auto conv_xy = [&](auto&& _1, auto&& _2) { return dy<D1FirstOrderCentered>(d1IntpCenterToCorner<1>(_1) * d1IntpCenterToCorner<0>(_2)); };
    auto conv_xz = [&](auto&& _1, auto&& _2) { return dz<D1FirstOrderCentered>(d1IntpCenterToCorner<2>(_1) * d1IntpCenterToCorner<0>(_2)); };
    auto conv_yx = [&](auto&& _1, auto&& _2) { return dx<D1FirstOrderCentered>(d1IntpCenterToCorner<1>(_1) * d1IntpCenterToCorner<0>(_2)); };
    auto conv_yy = [&](auto&& _1, auto&& _2) { return dy<D1FirstOrderCentered>(d1IntpCornerToCenter<1>(_1) * d1IntpCornerToCenter<1>(_2)); };
    auto conv_yz = [&](auto&& _1, auto&& _2) { return dz<D1FirstOrderCentered>(d1IntpCenterToCorner<2>(_1) * d1IntpCenterToCorner<1>(_2)); };
    auto conv_zx = [&](auto&& _1, auto&& _2) { return dx<D1FirstOrderCentered>(d1IntpCenterToCorner<2>(_1) * d1IntpCenterToCorner<0>(_2)); };
    auto conv_zy = [&](auto&& _1, auto&& _2) { return dy<D1FirstOrderCentered>(d1IntpCenterToCorner<2>(_1) * d1IntpCenterToCorner<1>(_2)); };
    auto conv_zz = [&](auto&& _1, auto&& _2) { return dz<D1FirstOrderCentered>(d1IntpCornerToCenter<2>(_1) * d1IntpCornerToCenter<2>(_2)); };
    auto laplace = [&](auto&& _1) { return d2x<D2SecondOrderCentered>(_1) + d2y<D2SecondOrderCentered>(_1) + d2z<D2SecondOrderCentered>(_1); };

    // solvers (LidDriven3D.cpp:91-103; tolerances from the command line so that both builds can be driven to the same solution)
    StructSolverParams<StructSolverType::GMRES> params;
    params.tol = tol;
    params.maxIter = 200;
    StructSolverParams<StructSolverType::PCG> poisson_params;
    poisson_params.tol = tol;
    poisson_params.maxIter = 200;
    StructSolverParams<StructSolverType::PFMG> p_params {.useZeroGuess = true, .relaxType = 1, .rapType = 0, .numPreRelax = 1, .numPostRelax = 1, .skipRelax = 0};
    p_params.tol = 1e-10;
    auto solver = PrecondStructSolver<StructSolverType::GMRES, StructSolverType::PFMG>(params, p_params);
    auto u_handler = makeEqnSolveHandler(
            [&](auto&& e) {
                return e / dt + conv_xx(u, e) + 0.5 * conv_xy(e, v) + 0.5 * conv_xz(e, w)
                       == nu * laplace(u) + 0.5 * nu * laplace(e) - (conv_xx(u, u) + conv_xy(u, v) + conv_xz(u, w)) - dx<D1FirstOrderCentered>(p);
            },
            du, solver);
    auto v_handler = makeEqnSolveHandler(
            [&](auto&& e) {
                return e / dt + conv_yy(v, e) + conv_yy(v, v) + conv_yx(u, v) + conv_yz(v, w) + 0.5 * conv_yx(u, e) + 0.5 * conv_yx(du, v) + 0.5 * conv_yz(e, w)
                       == nu * laplace(v) + 0.5 * nu * laplace(e) - dy<D1FirstOrderCentered>(p);
            },
            dv, solver);
    auto w_handler = makeEqnSolveHandler(
            [&](auto&& e) {
                return e / dt + 0.5 * conv_zx(u, e) + 0.5 * conv_zy(v, e) + conv_zz(w, e)
                       == nu * laplace(w) + 0.5 * nu * laplace(e) - conv_zx(u, w) - conv_zy(v, w) - conv_zz(w, w) - 0.5 * conv_zx(du, w) - 0.5 * conv_zy(dv, w)
                                  - dz<D1FirstOrderCentered>(p);
            },
            dw, solver);
    poisson_params.staticMat = true;
    poisson_params.pinValue = true;
    auto p_solver = PrecondStructSolver<StructSolverType::PCG, StructSolverType::PFMG>(poisson_params, p_params);
    auto p_handler = makeEqnSolveHandler(
            [&](auto&& e) {
                return laplace(e) * -1.0 == (dx<D1FirstOrderCentered>(du) + dy<D1FirstOrderCentered>(dv) + dz<D1FirstOrderCentered>(dw)) / -dt;
            },
            dp, p_solver);

    double t_mom = 0, t_exp = 0, t_poi = 0;
    int it_mom = 0, it_poi = 0;
    for (int i = 0; i < steps + 1; ++i) {// step 0 is the warm-up (allocations, solver set-up); it is part of the trajectory but not of the timing
        device_sync();
        double t0 = now();
        auto s1 = u_handler->solve();
        auto s2 = v_handler->solve();
        auto s3 = w_handler->solve();
        device_sync();
        double t1 = now();
        dv = dv - 0.5 * dt * conv_yz(v, dw);
        du = du - 0.5 * dt * conv_xy(u, dv) - 0.5 * dt * conv_xz(u, dw);
        u = u + du;
        v = v + dv;
        w = w + dw;
        device_sync();
        double t2 = now();
        auto s4 = p_handler->solve();
        device_sync();
        double t3 = now();
        u = u - dt * dx<D1FirstOrderCentered>(dp);
        v = v - dt * dy<D1FirstOrderCentered>(dp);
        w = w - dt * dz<D1FirstOrderCentered>(dp);
        p = p + dp;
        device_sync();
        double t4 = now();
        if (i > 0 || steps == 0) {
            t_mom += t1 - t0;
            t_exp += (t2 - t1) + (t4 - t3);
            t_poi += t3 - t2;
            it_mom += s1.niter + s2.niter + s3.niter;
            it_poi += s4.niter;
        }
    }
    const int timed = steps > 0 ? steps : 1;
    const long long cells = (long long) (n - 1) * (n - 1) * (nz - 1);
    if (getWorkerId() == 0)
        printf("{\"case\": \"tg3d\", \"n\": %d, \"nz\": %d, \"cells\": %lld, \"workers\": %d, \"threads\": %d, \"steps\": %d, \"tol\": %.1e, "
               "\"momentum_ms_per_step\": %.3f, \"explicit_ms_per_step\": %.3f, \"poisson_ms_per_step\": %.3f, \"momentum_iterations_per_step\": %.2f, "
               "\"poisson_iterations_per_step\": %.2f, \"explicit_sweeps_per_step\": 9}\n",
               n, nz, cells, getWorkerCount(), nt, steps, tol, 1e3 * t_mom / timed, 1e3 * t_exp / timed, 1e3 * t_poi / timed, (double) it_mom / timed,
               (double) it_poi / timed);
    if (!dump.empty()) {
        const std::string sfx = getWorkerCount() > 1 ? "." + std::to_string(getWorkerId()) : "";
        dump_field(dump + "_u.opfd" + sfx, u, false);
        dump_field(dump + "_v.opfd" + sfx, v, false);
        dump_field(dump + "_w.opfd" + sfx, w, false);
        dump_field(dump + "_p.opfd" + sfx, p, false);
    }
    return 0;
}
