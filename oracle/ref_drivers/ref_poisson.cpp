// oracle/ref_drivers/ref_poisson.cpp -- TEST INFRASTRUCTURE.
// BASELINE config C4's timed kernel on the UNMODIFIED reference: the pressure Poisson handler of
// examples/LidDriven/LidDriven2D.cpp:45-48,67-74 (d2x(e) + d2y(e) == b, Neumann on every side, pinValue, staticMat,
// GMRES + PFMG, tol 1e-10) on n x n nodes, i.e. (n-1)^2 pressure cells, with the manufactured right-hand side
// b = L_h(cos(2 pi x) cos(pi y)).  Prints one JSON line: milliseconds of the first solve (matrix assembly + HYPRE setup) and of
// the following ones, iterations, relative residual; --dump writes p.
//   ref_poisson --n N --solves S --threads T --tol 1e-10 --dump path
#include "ref_common.hpp"
using namespace OpFlow;
using namespace refdrv;

int main(int argc, char** argv) {
    EnvironmentGardian _env(&argc, &argv);
    const int n = atoi(arg(argc, argv, "--n", "257"));
    const int solves = atoi(arg(argc, argv, "--solves", "3"));
    const int nt = atoi(arg(argc, argv, "--threads", "1"));
    const char* dump = arg(argc, argv, "--dump", "");
    const double tol = atof(arg(argc, argv, "--tol", "1e-10"));
    set_threads(nt);
    using Mesh = CartesianMesh<Meta::int_<2>>;
    using Field = CartesianField<Real, Mesh>;
    auto m = MeshBuilder<Mesh>().newMesh(n, n).setMeshOfDim(0, 0., 1.).setMeshOfDim(1, 0., 1.).build();
    auto p = ExprBuilder<Field>().setMesh(m).setName("p").setBC(0, DimPos::start, BCType::Neum, 0.).setBC(0, DimPos::end, BCType::Neum, 0.)
                     .setBC(1, DimPos::start, BCType::Neum, 0.).setBC(1, DimPos::end, BCType::Neum, 0.).setExt(1)
                     .setLoc({LocOnMesh::Center, LocOnMesh::Center}).build();
    auto pt = p;
    auto b = p;
    pt.initBy([&](auto&& x) { return std::cos(2 * PI * x[0]) * std::cos(PI * x[1]); });
    b = d2x<D2SecondOrderCentered>(pt) + d2y<D2SecondOrderCentered>(pt);
    StructSolverParams<StructSolverType::GMRES> params;
    params.tol = tol;
    params.maxIter = 100;
    params.staticMat = true;
    params.pinValue = true;
    StructSolverParams<StructSolverType::PFMG> p_params {.useZeroGuess = true, .relaxType = 1, .rapType = 0, .numPreRelax = 1, .numPostRelax = 1, .skipRelax = 0};
    p_params.tol = 1e-10;
    auto solver = PrecondStructSolver<StructSolverType::GMRES, StructSolverType::PFMG>(params, p_params);
    auto handler = makeEqnSolveHandler([&](auto&& e) { return d2x<D2SecondOrderCentered>(e) + d2y<D2SecondOrderCentered>(e) == b; }, p, solver);
    std::vector<double> ms;
    EqnSolveState st;
    for (int i = 0; i < solves; ++i) {
        p = 0.;
        const double t0 = now();
        st = handler->solve();
        ms.push_back((now() - t0) * 1e3);
    }
    double rest = 0;
    for (size_t i = 1; i < ms.size(); ++i) rest += ms[i];
    if (ms.size() > 1) rest /= (double) (ms.size() - 1);
    else rest = ms[0];
    printf("{\"case\": \"poisson2d\", \"n\": %d, \"cells\": %lld, \"threads\": %d, \"solves\": %d, \"first_solve_ms\": %.3f, \"ms_per_solve\": %.3f, "
           "\"niter\": %d, \"relerr\": %.3e}\n",
           n, (long long) (n - 1) * (n - 1), nt, solves, ms[0], rest, st.niter, st.relerr);
    if (dump && *dump) dump_field(dump, p, false);
    return 0;
}
