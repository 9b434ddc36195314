// oracle/ref_drivers/ref_ld2d.cpp -- TEST INFRASTRUCTURE.
// The lid-driven cavity of examples/LidDriven/LidDriven2D.cpp:10-96 with the mesh size, the step count and the solver tolerance on
// the command line (the example hard-codes n = 65 and 1000 steps): MAC-staggered u, v, p on [0,1]^2, semi-implicit convection /
// diffusion momentum solves for du, dv (GMRES + PFMG), explicit cross-term correction, pressure Poisson solve (GMRES + PFMG,
// staticMat + pinValue), projection.  Purpose: parity and per-step times of the variable-coefficient (momentum) solves at sizes
// where a diagonal preconditioner is not enough (nu dt / h^2 >> 1).
// Compiled twice from this one file: against the unmodified reference (oracle/build_ref.sh -> oracle/_ref/bin/ref_ld2d) and against
// the B200 front-end (tests/frontend -> fe_ld2d).
//   ref_ld2d --n N --steps S --threads T --tol 1e-10 --dump prefix [--stride K]
#include "ref_common.hpp"
using namespace OpFlow;
using namespace refdrv;

static void device_sync() {
#ifdef OPFLOW_B200
    opf_synchronize();
#endif
}

// every K-th value of the local range along each axis, OPFD header describing the sampled index box
template <typename F>
static void dump_sampled(const std::string& path, const F& f, int K) {
    FILE* fp = fopen(path.c_str(), "wb");
    if (!fp) { perror(path.c_str()); exit(2); }
    auto r = f.localRange;
    int32_t d = 2, s[2] = {0, 0}, e[2];
    for (int i = 0; i < 2; ++i) e[i] = (r.end[i] - r.start[i] + K - 1) / K;
    fwrite("OPFD", 1, 4, fp);
    fwrite(&d, 4, 1, fp);
    fwrite(s, 4, 2, fp);
    fwrite(e, 4, 2, fp);
    std::vector<double> buf;
    for (int j = 0; j < e[1]; ++j)
        for (int i = 0; i < e[0]; ++i) buf.push_back(f.evalAt(DS::MDIndex<2> {r.start[0] + i * K, r.start[1] + j * K}));
    fwrite(buf.data(), 8, buf.size(), fp);
    fclose(fp);
}

int main(int argc, char** argv) {
    EnvironmentGardian _env(&argc, &argv);
    using Mesh = CartesianMesh<Meta::int_<2>>;
    using Field = CartesianField<Real, Mesh>;
    const int n = atoi(arg(argc, argv, "--n", "65"));
    const int steps = atoi(arg(argc, argv, "--steps", "5"));
    const int nt = atoi(arg(argc, argv, "--threads", "1"));
    const double tol = atof(arg(argc, argv, "--tol", "1e-10"));
    const int stride = atoi(arg(argc, argv, "--stride", "1"));
    const std::string dump = arg(argc, argv, "--dump", "");
    set_threads(nt);

    auto mesh = MeshBuilder<Mesh>().newMesh(n, n).setMeshOfDim(0, 0., 1.).setMeshOfDim(1, 0., 1.).build();
    auto builder = ExprBuilder<Field>().setMesh(mesh).setExt(1)
                           .setBC(0, DimPos::start, BCType::Dirc, 0.).setBC(0, DimPos::end, BCType::Dirc, 0.)
                           .setBC(1, DimPos::start, BCType::Dirc, 0.).setBC(1, DimPos::end, BCType::Dirc, 0.);
    auto u = builder.setName("u").setBC(1, DimPos::end, BCType::Dirc, 1.).setLoc({LocOnMesh::Corner, LocOnMesh::Center}).build();
    auto du = builder.setName("du").setBC(1, DimPos::end, BCType::Dirc, 0.).build();
    auto v = builder.setName("v").setBC(1, DimPos::end, BCType::Dirc, 0.).setLoc({LocOnMesh::Center, LocOnMesh::Corner}).build();
    auto dv = v;
    dv.name = "dv";
    auto p = builder.setName("p")
                     .setBC(0, DimPos::start, BCType::Neum, 0.).setBC(0, DimPos::end, BCType::Neum, 0.)
                     .setBC(1, DimPos::start, BCType::Neum, 0.).setBC(1, DimPos::end, BCType::Neum, 0.)
                     .setLoc({LocOnMesh::Center, LocOnMesh::Center}).build();
    auto dp = p;
    dp.name = "dp";
    u = 0; du = 0; v = 0; dv = 0; p = 0; dp = 0;

    auto conv_xx = [&](auto&& _1, auto&& _2) { return dx<D1FirstOrderCentered>(d1IntpCornerToCenter<0>(_1) * d1IntpCornerToCenter<0>(_2)); };
    auto conv_xy = [&](auto&& _1, auto&& _2) { return dy<D1FirstOrderCentered>(d1IntpCenterToCorner<1>(_1) * d1IntpCenterToCorner<0>(_2)); };
    auto conv_yx = [&](auto&& _1, auto&& _2) { return dx<D1FirstOrderCentered>(d1IntpCenterToCorner<1>(_1) * d1IntpCenterToCorner<0>(_2)); };
    auto conv_yy = [&](auto&& _1, auto&& _2) { return dy<D1FirstOrderCentered>(d1IntpCornerToCenter<1>(_1) * d1IntpCornerToCenter<1>(_2)); };
    auto laplace = [&](auto&& _1) { return d2x<D2SecondOrderCentered>(_1) + d2y<D2SecondOrderCentered>(_1); };

    const Real dt = 0.5e-2, nu = 1.0e-2;
    StructSolverParams<StructSolverType::GMRES> params;
    params.tol = tol;
    params.maxIter = 100;
    StructSolverParams<StructSolverType::GMRES> poisson_params = params;
    StructSolverParams<StructSolverType::PFMG> p_params {.useZeroGuess = true, .relaxType = 1, .rapType = 0, .numPreRelax = 1, .numPostRelax = 1, .skipRelax = 0};
    p_params.tol = 1e-10;
    auto solver = PrecondStructSolver<StructSolverType::GMRES, StructSolverType::PFMG>(params, p_params);
    auto u_handler = makeEqnSolveHandler(
            [&](auto&& e) {
                return e / dt + conv_xx(u, e) + 0.5 * conv_xy(e, v)
                       == nu * laplace(u) + 0.5 * nu * laplace(e) - (conv_xx(u, u) + conv_xy(u, v)) - dx<D1FirstOrderCentered>(p);
            },
            du, solver);
    auto v_handler = makeEqnSolveHandler(
            [&](auto&& e) {
                return e / dt + conv_yy(v, e) + conv_yy(v, v) + conv_yx(u, v) + 0.5 * conv_yx(u, e) + 0.5 * conv_yx(du, v)
                       == nu * laplace(v) + 0.5 * nu * laplace(e) - dy<D1FirstOrderCentered>(p);
            },
            dv, solver);
    poisson_params.staticMat = true;
    poisson_params.pinValue = true;
    auto p_solver = PrecondStructSolver<StructSolverType::GMRES, StructSolverType::PFMG>(poisson_params, p_params);
    auto p_handler = makeEqnSolveHandler(
            [&](auto&& e) { return laplace(e) == (dx<D1FirstOrderCentered>(du) + dy<D1FirstOrderCentered>(dv)) / dt; }, dp, p_solver);

    double t_mom = 0, t_exp = 0, t_poi = 0;
    int it_mom = 0, it_poi = 0;
    for (int i = 0; i < steps + 1; ++i) {// step 0 is the warm-up (allocations, solver set-up); part of the trajectory, not of the timing
        device_sync();
        const double t0 = now();
        auto s1 = u_handler->solve();
        auto s2 = v_handler->solve();
        device_sync();
        const double t1 = now();
        du = du - 0.5 * dt * conv_xy(u, dv);
        u = u + du;
        v = v + dv;
        device_sync();
        const double t2 = now();
        auto s3 = p_handler->solve();
        device_sync();
        const double t3 = now();
        u = u - dt * dx<D1FirstOrderCentered>(dp);
        v = v - dt * dy<D1FirstOrderCentered>(dp);
        p = p + dp;
        device_sync();
        const double t4 = now();
        if (i > 0 || steps == 0) {
            t_mom += t1 - t0;
            t_exp += (t2 - t1) + (t4 - t3);
            t_poi += t3 - t2;
            it_mom += s1.niter + s2.niter;
            it_poi += s3.niter;
        }
    }
    const int timed = steps > 0 ? steps : 1;
    printf("{\"case\": \"ld2d\", \"n\": %d, \"cells\": %lld, \"threads\": %d, \"steps\": %d, \"tol\": %.1e, \"momentum_ms_per_step\": %.3f, "
           "\"explicit_ms_per_step\": %.3f, \"poisson_ms_per_step\": %.3f, \"momentum_iterations_per_step\": %.2f, \"poisson_iterations_per_step\": %.2f}\n",
           n, (long long) (n - 1) * (n - 1), nt, steps, tol, 1e3 * t_mom / timed, 1e3 * t_exp / timed, 1e3 * t_poi / timed, (double) it_mom / timed,
           (double) it_poi / timed);
    if (!dump.empty()) {
        dump_sampled(dump + "_u.opfd", u, stride);
        dump_sampled(dump + "_v.opfd", v, stride);
        dump_sampled(dump + "_p.opfd", p, stride);
    }
    return 0;
}
