// engine_solver.cu -- the implicit path, matrix-free on the device.
//
// Reference: HYPREEqnSolveHandler (src/Core/Equation/HYPREEqnSolveHandler.hpp:50-231) evaluates the user's equation on a
// symbolic StencilField, pushes every matrix row into a HYPRE Struct matrix through host SetValues calls and re-creates /
// re-sets-up the HYPRE solver on every solve().  Here nothing is assembled: the operator of  lhs(e) == rhs  is applied by
// running the *same device functor* the explicit path uses, on a work vector whose ghost cells are filled with the
// target's boundary conditions made homogeneous (A.p = lhs(G_h(p)) - lhs(G_h(0))), and b = rhs - lhs(G(0)) carries the
// boundary data and any unknown-free terms of lhs.
// Krylov: PCG (HYPRE_StructPCG*, StructSolverPCG.hpp:25-96) and BiCGSTAB (StructSolverBiCGSTAB.hpp) -- the latter also
// serves GMRES-family requests.  Preconditioner / stand-alone solver: geometric multigrid V-cycle (PFMG's role,
// StructSolverPFMG.hpp:36-110) with weighted-Jacobi relaxation (one fused kernel per sweep where the expression
// x + w*dinv*(b - lhs(x)) is compiled in), full coarsening, operators re-discretised on the coarse meshes by the same
// functor, transfer operators with R = (1/2) P^T per axis (node- or cell-centred per LocOnMesh), mean projection on singular
// levels; or plain Jacobi (StructSolverJacobi.hpp).  The V-cycle runs directly on the Krylov vectors and is replayed as a
// CUDA graph (keyed by the ping-pong state of the level iterates).
// Decomposed targets: every level keeps the target's block decomposition (boundaries halved, halo exchange inside
// updatePadding, NCCL allreduce for dot products / projections, the pinned cell handled by its owner) until blocks get
// thinner than two halos; below that the hierarchy continues replicated on every rank behind one allreduce.
#include "engine.hpp"
#include <algorithm>
#include <cmath>

namespace opfe {
    int signature_radius(const char* sig);// engine_expr.cu
    int signature_unknown_taps(const char* sig, unsigned mask, std::vector<std::array<int, 3>>& out);
    int reduce_sum_device(opf_field_s* f, const Range& r, double** dev_result);
    int dot_device(opf_field_s* a, opf_field_s* b, const Range& r, double* slot);
}
using namespace opfe;

namespace {
    struct XferParams {
        double* fine;
        long long fs1, fs2;
        double* coarse;
        long long cs1, cs2;
        int dim;
        int flo[3], fhi[3];// fine writable range of the whole (global) field: validity + wall logic
        int clo[3], chi[3];// coarse writable range of the whole (global) field
        int wlo[3], whi[3];// the box this launch writes (this rank's part of the coarse / fine level)
        // decomposed levels keep their periodic images in the halo (filled by the exchange): indices are then NOT wrapped, the
        // ghost cell is read instead -- the wrapped cell may live on another rank
        int fhalo[3], chalo[3];
        int f0[3], c0[3];  // accessible.start of fine / coarse (index origin of the 2:1 map)
        int center[3], periodic[3], fper[3], cper[3];
        int bclo[3], bchi[3];// opf_bctype of the unknown per axis side (cell-centred prolongation at the walls)
    };

    __device__ __forceinline__ bool in_box(const int* g, const int* lo, const int* hi) {
        return g[0] >= lo[0] && g[0] < hi[0] && g[1] >= lo[1] && g[1] < hi[1] && g[2] >= lo[2] && g[2] < hi[2];
    }
    // same, but axes flagged in `halo` accept any index (their out-of-range cells are valid periodic images in the halo)
    __device__ __forceinline__ bool in_box_halo(const int* g, const int* lo, const int* hi, const int* halo) {
        return (halo[0] || (g[0] >= lo[0] && g[0] < hi[0])) && (halo[1] || (g[1] >= lo[1] && g[1] < hi[1])) && (halo[2] || (g[2] >= lo[2] && g[2] < hi[2]));
    }

    // per-axis prolongation weight of coarse cell I for fine cell i (cell-centred axes), shared by both transfer kernels so
    // that R = (1/2) P^T per axis exactly (a symmetric V-cycle is what lets it precondition CG)
    __host__ __device__ __forceinline__ double cc_weight(const XferParams& p, int d, int i, int I) {
        const int r = i - p.f0[d];
        const int par = p.c0[d] + (r >> 1);
        int nb = (r & 1) ? par + 1 : par - 1;
        bool wall = false;
        if (p.periodic[d]) {
            if (!p.chalo[d]) {
                if (nb < p.c0[d]) nb += p.cper[d];
                else if (nb >= p.c0[d] + p.cper[d])
                    nb -= p.cper[d];
            }
        } else if (nb < p.clo[d] || nb >= p.chi[d]) {
            wall = true;
        }
        if (I == par) {
            if (!wall) return 0.75;
            const int bt = nb < p.clo[d] ? p.bclo[d] : p.bchi[d];
            return (bt == OPF_BC_DIRC || bt == OPF_BC_ASYMM) ? 0.5 : 1.0;
        }
        if (!wall && I == nb) return 0.25;
        return 0.0;
    }

    // coarse = R fine with R = (1/2) P^T per axis: node-centred (1/4, 1/2, 1/4) full weighting; cell-centred
    // (1/8, 3/8, 3/8, 1/8) over fine cells 2I-1 .. 2I+2 (wall-adjusted like the prolongation).
    __global__ void __launch_bounds__(256) restrict_kernel(const XferParams p) {
        // one thread per coarse cell; rows / planes come from the grid (no 64-bit divisions), axis 0 is coalesced
        {
            const int x0 = blockIdx.x * blockDim.x + threadIdx.x;
            if (x0 >= p.whi[0] - p.wlo[0]) return;
            int I[3] = {p.wlo[0] + x0, p.wlo[1] + (int) blockIdx.y, p.wlo[2] + (int) blockIdx.z};
            int idx[3][4], cnt[3];
            double wgt[3][4];
            for (int d = 0; d < 3; ++d) {
                if (d >= p.dim) {
                    idx[d][0] = 0, cnt[d] = 1, wgt[d][0] = 1.0;
                } else if (p.center[d]) {
                    cnt[d] = 4;
                    for (int a = 0; a < 4; ++a) {
                        int i = p.f0[d] + 2 * (I[d] - p.c0[d]) - 1 + a;
                        if (p.periodic[d] && !p.fhalo[d]) {
                            if (i < p.f0[d]) i += p.fper[d];
                            else if (i >= p.f0[d] + p.fper[d])
                                i -= p.fper[d];
                        }
                        idx[d][a] = i;
                        wgt[d][a] = ((i >= p.flo[d] && i < p.fhi[d]) || (p.periodic[d] && p.fhalo[d])) ? 0.5 * cc_weight(p, d, i, I[d]) : 0.0;
                    }
                } else {
                    cnt[d] = 3;
                    for (int a = 0; a < 3; ++a) {
                        int i = p.f0[d] + 2 * (I[d] - p.c0[d]) - 1 + a;
                        if (p.periodic[d] && !p.fhalo[d]) {
                            if (i < p.f0[d]) i += p.fper[d];
                            else if (i >= p.f0[d] + p.fper[d])
                                i -= p.fper[d];
                        }
                        idx[d][a] = i;
                        wgt[d][a] = a == 1 ? 0.5 : 0.25;
                    }
                }
            }
            double acc = 0.0;
            for (int c = 0; c < cnt[2]; ++c)
                for (int b = 0; b < cnt[1]; ++b)
                    for (int a = 0; a < cnt[0]; ++a) {
                        const int g[3] = {idx[0][a], idx[1][b], idx[2][c]};
                        const double w = wgt[0][a] * wgt[1][b] * wgt[2][c];
                        if (w == 0.0 || !in_box_halo(g, p.flo, p.fhi, p.fhalo)) continue;
                        acc += w * p.fine[(long long) g[0] + (long long) g[1] * p.fs1 + (long long) g[2] * p.fs2];
                    }
            p.coarse[(long long) I[0] + (long long) I[1] * p.cs1 + (long long) I[2] * p.cs2] = acc;
        }
    }

    // fine += P coarse.  Linear interpolation per axis: node-centred (Corner) weights (1) / (1/2, 1/2); cell-centred (Center)
    // weights (3/4 parent, 1/4 neighbouring coarse cell).  A neighbour beyond a wall is the homogeneous ghost value of the
    // coarse correction: +parent for Neumann/Symm, -parent for Dirichlet/ASymm.
    __global__ void __launch_bounds__(256) prolong_kernel(const XferParams p) {
        {
            const int x0 = blockIdx.x * blockDim.x + threadIdx.x;
            if (x0 >= p.whi[0] - p.wlo[0]) return;
            int g[3] = {p.wlo[0] + x0, p.wlo[1] + (int) blockIdx.y, p.wlo[2] + (int) blockIdx.z};
            int idx[3][2], cnt[3];
            double wgt[3][2];
            for (int d = 0; d < 3; ++d) {
                if (d >= p.dim) {
                    idx[d][0] = 0, cnt[d] = 1, wgt[d][0] = 1.0;
                    continue;
                }
                const int r = g[d] - p.f0[d];
                const int par = p.c0[d] + (r >> 1);
                if (p.center[d]) {
                    int nb = (r & 1) ? par + 1 : par - 1;
                    if (p.periodic[d] && !p.chalo[d]) {
                        if (nb < p.c0[d]) nb += p.cper[d];
                        else if (nb >= p.c0[d] + p.cper[d])
                            nb -= p.cper[d];
                    }
                    idx[d][0] = par, wgt[d][0] = cc_weight(p, d, g[d], par);
                    idx[d][1] = nb, wgt[d][1] = cc_weight(p, d, g[d], nb), cnt[d] = wgt[d][1] != 0.0 ? 2 : 1;
                } else if ((r & 1) == 0) {
                    idx[d][0] = par, cnt[d] = 1, wgt[d][0] = 1.0;
                } else {
                    idx[d][0] = par, idx[d][1] = par + 1, cnt[d] = 2, wgt[d][0] = wgt[d][1] = 0.5;
                    if (p.periodic[d] && !p.chalo[d] && idx[d][1] >= p.c0[d] + p.cper[d]) idx[d][1] -= p.cper[d];
                }
            }
            double acc = 0.0;
            for (int c = 0; c < cnt[2]; ++c)
                for (int b = 0; b < cnt[1]; ++b)
                    for (int a = 0; a < cnt[0]; ++a) {
                        const int I[3] = {idx[0][a], idx[1][b], idx[2][c]};
                        if (!in_box_halo(I, p.clo, p.chi, p.chalo)) continue;
                        acc += wgt[0][a] * wgt[1][b] * wgt[2][c] * p.coarse[(long long) I[0] + (long long) I[1] * p.cs1 + (long long) I[2] * p.cs2];
                    }
            p.fine[(long long) g[0] + (long long) g[1] * p.fs1 + (long long) g[2] * p.fs2] += acc;
        }
    }

    // ---- table-driven transfer kernels (2-D / 3-D).  The per-axis index / weight logic of restrict_kernel and prolong_kernel above
    // (parents, neighbours, wall and periodic handling) is evaluated ONCE per level on the host into small per-axis tables; the
    // kernels then do nothing but table look-ups, 4 / 8 (prolong) or up to 16 / 64 (restrict) cached loads and the multiply-adds:
    // no branches, no local-memory index arrays.  ncu (profiles/r2c_c4_launches.csv): level-0 prolongation 239 us -> HBM-bound.
    struct AxisTab {
        const int* pidx;   // [pn][2] coarse indices feeding fine index plo + q
        const double* pw;  // [pn][2] their weights (0 where the tap does not exist)
        const int* ridx;   // [rn][4] fine indices feeding coarse index rlo + q
        const double* rw;  // [rn][4] their weights
        int plo, pn, rlo, rn, rcnt;
    };
    struct XferTab {
        AxisTab ax[3];
    };
    inline void host_prolong_axis(const XferParams& p, int d, int g, int idx[2], double w[2]) {
        idx[0] = idx[1] = 0, w[0] = 1.0, w[1] = 0.0;
        if (d >= p.dim) return;
        const int r = g - p.f0[d];
        const int par = p.c0[d] + (r >> 1);
        if (p.center[d]) {
            int nb = (r & 1) ? par + 1 : par - 1;
            if (p.periodic[d] && !p.chalo[d]) {
                if (nb < p.c0[d]) nb += p.cper[d];
                else if (nb >= p.c0[d] + p.cper[d])
                    nb -= p.cper[d];
            }
            idx[0] = par, w[0] = cc_weight(p, d, g, par);
            idx[1] = nb, w[1] = cc_weight(p, d, g, nb);
        } else if ((r & 1) == 0) {
            idx[0] = idx[1] = par, w[0] = 1.0, w[1] = 0.0;
        } else {
            idx[0] = par, idx[1] = par + 1, w[0] = w[1] = 0.5;
            if (p.periodic[d] && !p.chalo[d] && idx[1] >= p.c0[d] + p.cper[d]) idx[1] -= p.cper[d];
        }
        for (int a = 0; a < 2; ++a)
            if (!p.chalo[d] && (idx[a] < p.clo[d] || idx[a] >= p.chi[d])) {// outside the coarse box: the tap does not exist
                w[a] = 0.0;
                idx[a] = idx[a] < p.clo[d] ? p.clo[d] : p.chi[d] - 1;
            }
    }
    inline int host_restrict_axis(const XferParams& p, int d, int I, int idx[4], double w[4]) {
        for (int a = 0; a < 4; ++a) idx[a] = 0, w[a] = 0.0;
        if (d >= p.dim) {
            w[0] = 1.0;
            return 1;
        }
        const int cnt = p.center[d] ? 4 : 3;
        for (int a = 0; a < cnt; ++a) {
            int i = p.f0[d] + 2 * (I - p.c0[d]) - 1 + a;
            if (p.periodic[d] && !p.fhalo[d]) {
                if (i < p.f0[d]) i += p.fper[d];
                else if (i >= p.f0[d] + p.fper[d])
                    i -= p.fper[d];
            }
            idx[a] = i;
            if (p.center[d]) w[a] = ((i >= p.flo[d] && i < p.fhi[d]) || (p.periodic[d] && p.fhalo[d])) ? 0.5 * cc_weight(p, d, i, I) : 0.0;
            else
                w[a] = a == 1 ? 0.5 : 0.25;
            if (!p.fhalo[d] && (i < p.flo[d] || i >= p.fhi[d])) {
                w[a] = 0.0;
                idx[a] = i < p.flo[d] ? p.flo[d] : p.fhi[d] - 1;
            }
        }
        for (int a = cnt; a < 4; ++a) idx[a] = idx[cnt - 1];// unused tap: weight 0, a valid address
        return cnt;
    }

    template <int DIM>
    __global__ void __launch_bounds__(256) prolong_fast_kernel(double* __restrict__ fine, long long fs1, long long fs2, const double* __restrict__ coarse,
                                                               long long cs1, long long cs2, opf::LaunchRange w, const __grid_constant__ XferTab t) {
        const int i = w.lo[0] + blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= w.hi[0]) return;
        const int j = w.lo[1] + (int) blockIdx.y, k = w.lo[2] + (int) blockIdx.z;
        const int qx = 2 * (i - t.ax[0].plo), qy = 2 * (j - t.ax[1].plo);
        const int ix0 = t.ax[0].pidx[qx], ix1 = t.ax[0].pidx[qx + 1];
        const double wx0 = t.ax[0].pw[qx], wx1 = t.ax[0].pw[qx + 1];
        const long long jy0 = (long long) t.ax[1].pidx[qy] * cs1, jy1 = (long long) t.ax[1].pidx[qy + 1] * cs1;
        const double wy0 = t.ax[1].pw[qy], wy1 = t.ax[1].pw[qy + 1];
        double acc;
        if constexpr (DIM == 2) {
            acc = wy0 * (wx0 * __ldg(coarse + ix0 + jy0) + wx1 * __ldg(coarse + ix1 + jy0)) + wy1 * (wx0 * __ldg(coarse + ix0 + jy1) + wx1 * __ldg(coarse + ix1 + jy1));
        } else {
            const int qz = 2 * (k - t.ax[2].plo);
            const long long kz0 = (long long) t.ax[2].pidx[qz] * cs2, kz1 = (long long) t.ax[2].pidx[qz + 1] * cs2;
            const double wz0 = t.ax[2].pw[qz], wz1 = t.ax[2].pw[qz + 1];
            const double a0 = wy0 * (wx0 * __ldg(coarse + ix0 + jy0 + kz0) + wx1 * __ldg(coarse + ix1 + jy0 + kz0))
                              + wy1 * (wx0 * __ldg(coarse + ix0 + jy1 + kz0) + wx1 * __ldg(coarse + ix1 + jy1 + kz0));
            const double a1 = wy0 * (wx0 * __ldg(coarse + ix0 + jy0 + kz1) + wx1 * __ldg(coarse + ix1 + jy0 + kz1))
                              + wy1 * (wx0 * __ldg(coarse + ix0 + jy1 + kz1) + wx1 * __ldg(coarse + ix1 + jy1 + kz1));
            acc = wz0 * a0 + wz1 * a1;
        }
        const long long o = (long long) i + (long long) j * fs1 + (long long) k * fs2;
        fine[o] += acc;
    }
    template <int DIM>
    __global__ void __launch_bounds__(256) restrict_fast_kernel(const double* __restrict__ fine, long long fs1, long long fs2, double* __restrict__ coarse,
                                                                long long cs1, long long cs2, opf::LaunchRange w, const __grid_constant__ XferTab t) {
        const int I = w.lo[0] + blockIdx.x * blockDim.x + threadIdx.x;
        if (I >= w.hi[0]) return;
        const int J = w.lo[1] + (int) blockIdx.y, K = w.lo[2] + (int) blockIdx.z;
        const int qx = 4 * (I - t.ax[0].rlo), qy = 4 * (J - t.ax[1].rlo);
        int ix[4];
        double wx[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) ix[a] = t.ax[0].ridx[qx + a], wx[a] = t.ax[0].rw[qx + a];
        double acc = 0.0;
        if constexpr (DIM == 2) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const double wy = t.ax[1].rw[qy + b];
                const double* row = fine + (long long) t.ax[1].ridx[qy + b] * fs1;
                acc += wy * (wx[0] * __ldg(row + ix[0]) + wx[1] * __ldg(row + ix[1]) + wx[2] * __ldg(row + ix[2]) + wx[3] * __ldg(row + ix[3]));
            }
        } else {
            const int qz = 4 * (K - t.ax[2].rlo);
            for (int c = 0; c < t.ax[2].rcnt; ++c) {
                const double wz = t.ax[2].rw[qz + c];
                const double* pl = fine + (long long) t.ax[2].ridx[qz + c] * fs2;
                double accp = 0.0;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const double wy = t.ax[1].rw[qy + b];
                    const double* row = pl + (long long) t.ax[1].ridx[qy + b] * fs1;
                    accp += wy * (wx[0] * __ldg(row + ix[0]) + wx[1] * __ldg(row + ix[1]) + wx[2] * __ldg(row + ix[2]) + wx[3] * __ldg(row + ix[3]));
                }
                acc += wz * accp;
            }
        }
        coarse[(long long) I + (long long) J * cs1 + (long long) K * cs2] = acc;
    }

    // probe vector for the diagonal: 1 on the cells whose index is congruent to `col` modulo m on every axis
    // colour of a cell: cube colouring (i mod m0, j mod m1, k mod m2) == (c0, c1, c2), or -- lat[3] > 0 -- the lattice colouring
    // (lat[0] i + lat[1] j + lat[2] k) mod lat[3] == c0
    struct Mod3 {
        int m[3];
        int lat[4];
    };
    __device__ __forceinline__ bool colour_on(const Mod3& mm, int dim, int i, int j, int k, int c0, int c1, int c2) {
        if (mm.lat[3] > 0) {
            long long v = (long long) mm.lat[0] * i + (dim >= 2 ? (long long) mm.lat[1] * j : 0) + (dim >= 3 ? (long long) mm.lat[2] * k : 0);
            v %= mm.lat[3];
            if (v < 0) v += mm.lat[3];
            return (int) v == c0;
        }
        return ((i % mm.m[0] + mm.m[0]) % mm.m[0] == c0) && (dim < 2 || (j % mm.m[1] + mm.m[1]) % mm.m[1] == c1) && (dim < 3 || (k % mm.m[2] + mm.m[2]) % mm.m[2] == c2);
    }
    __global__ void __launch_bounds__(256) color_fill_kernel(double* u, long long s1, long long s2, opf::LaunchRange r, int dim, Mod3 mm, int c0, int c1,
                                                             int c2) {
        {
            const int x0 = blockIdx.x * blockDim.x + threadIdx.x;
            if (x0 >= r.hi[0] - r.lo[0]) return;
            const int i = r.lo[0] + x0, j = r.lo[1] + (int) blockIdx.y, k = r.lo[2] + (int) blockIdx.z;
            const bool on = colour_on(mm, dim, i, j, k, c0, c1, c2);
            u[(long long) i + (long long) j * s1 + (long long) k * s2] = on ? 1.0 : 0.0;
        }
    }
    __global__ void __launch_bounds__(256) color_recip_kernel(const double* q, double* dinv, long long s1, long long s2, opf::LaunchRange r, int dim,
                                                              Mod3 mm, int c0, int c1, int c2) {
        {
            const int x0 = blockIdx.x * blockDim.x + threadIdx.x;
            if (x0 >= r.hi[0] - r.lo[0]) return;
            const int i = r.lo[0] + x0, j = r.lo[1] + (int) blockIdx.y, k = r.lo[2] + (int) blockIdx.z;
            const bool on = colour_on(mm, dim, i, j, k, c0, c1, c2);
            if (on) {
                const long long o = (long long) i + (long long) j * s1 + (long long) k * s2;
                const double d = q[o];
                dinv[o] = d != 0.0 ? 1.0 / d : 0.0;
            }
        }
    }
    // u -= sum[0] / n over a box: projects a coarse right-hand side onto the range of a singular (all-Neumann / periodic)
    // operator without a host round trip (sum[0] was produced by the reduction kernels on the same stream)
    __global__ void __launch_bounds__(256) sub_mean_kernel(double* u, long long s1, long long s2, opf::LaunchRange r, const double* sum, double inv_n) {
        const double m = sum[0] * inv_n;
        const int x0 = blockIdx.x * blockDim.x + threadIdx.x;
        if (x0 >= r.hi[0] - r.lo[0]) return;
        const long long o = (long long) (r.lo[0] + x0) + (long long) (r.lo[1] + (int) blockIdx.y) * s1 + (long long) (r.lo[2] + (int) blockIdx.z) * s2;
        u[o] -= m;
    }
    // coefficient field of a coarse level = its fine-level parent sampled at the coarse unknown positions: a Corner axis injects the
    // coinciding node (coarse node i = fine node 2i), a Center axis averages the two fine cells under the coarse cell
    struct CoefMap {
        int f0[3], c0[3];// first mesh index of the fine / coarse level
        int avg[3];      // 1: Center axis (two children), 0: Corner axis (injection) or unused axis
    };
    __global__ void __launch_bounds__(256) coef_restrict_kernel(const double* __restrict__ fine, long long fs1, long long fs2, double* __restrict__ coarse,
                                                                long long cs1, long long cs2, opf::LaunchRange r, CoefMap m) {
        const int x0 = blockIdx.x * blockDim.x + threadIdx.x;
        if (x0 >= r.hi[0] - r.lo[0]) return;
        const int i = r.lo[0] + x0, j = r.lo[1] + (int) blockIdx.y, k = r.lo[2] + (int) blockIdx.z;
        const int fi = m.f0[0] + 2 * (i - m.c0[0]), fj = m.f0[1] + 2 * (j - m.c0[1]), fk = m.f0[2] + 2 * (k - m.c0[2]);
        double acc = 0.0;
        for (int c = 0; c <= m.avg[2]; ++c)
            for (int b = 0; b <= m.avg[1]; ++b)
                for (int a = 0; a <= m.avg[0]; ++a) acc += fine[(long long) (fi + a) + (long long) (fj + b) * fs1 + (long long) (fk + c) * fs2];
        const int cnt = (1 + m.avg[0]) * (1 + m.avg[1]) * (1 + m.avg[2]);
        coarse[(long long) i + (long long) j * cs1 + (long long) k * cs2] = acc / (double) cnt;
    }
    __global__ void set_cell_kernel(double* u, long long off, double v) { u[off] = v; }
    __global__ void copy_cell_kernel(double* dst, const double* src, long long off) { dst[off] = src[off]; }

    // ------------------------------------------------------------------------------------------------ device-resident PCG
    // The Krylov scalars never visit the host: dot products fold into ks[], the update kernels form alpha / beta from them, and a
    // control kernel decides whether the iteration continues -- through the condition of a CUDA-graph WHILE node when the whole
    // iteration is captured (one graph launch per solve), through a mapped flag read once per iteration otherwise.
    enum { KS_RZ = 0, KS_PQ, KS_RR, KS_RZN, KS_BNORM, KS_TOL, KS_COUNT = 8 };
    struct PcgCtl {
        int iters, maxit, cont, breakdown;
        double rel;
    };
    // work item of the streaming Krylov kernels: (row, 1024-element segment of axis 0); a block takes items round-robin, its 256 threads
    // take 2 consecutive cells each per trip and two trips are in flight (4 x 128-bit loads per operand per thread)
    constexpr int KSEG = 1024;
    struct RowIter {
        long long nseg, items;
        int n0, n1;
        __device__ __forceinline__ RowIter(const opf::LaunchRange& w) {
            n0 = w.hi[0] - w.lo[0], n1 = w.hi[1] - w.lo[1];
            nseg = (n0 + KSEG - 1) / KSEG;
            items = nseg * n1 * (long long) (w.hi[2] - w.lo[2]);
        }
        // -> offset of the segment's first cell, number of cells in it
        __device__ __forceinline__ long long locate(const opf::LaunchRange& w, long long it, long long s1, long long s2, int& len) const {
            const long long seg = it % nseg, row = it / nseg;
            const int j = w.lo[1] + (int) (row % n1), k = w.lo[2] + (int) (row / n1);
            const int b = (int) seg * KSEG;
            len = min(KSEG, n0 - b);
            return (long long) (w.lo[0] + b) + (long long) j * s1 + (long long) k * s2;
        }
    };
    // x += alpha p ; r -= alpha q ; partial sums of r.r          (48 B per cell instead of 24 + 24 + 8 in three passes)
    __global__ void __launch_bounds__(256) pcg_update_kernel(double* __restrict__ x, const double* __restrict__ p, double* __restrict__ r,
                                                             const double* __restrict__ q, long long s1, long long s2, opf::LaunchRange w,
                                                             const double* __restrict__ ks, double* __restrict__ partials, int vec) {
        const double pq = ks[KS_PQ];
        const double alpha = pq != 0.0 ? ks[KS_RZ] / pq : 0.0;
        const RowIter ri(w);
        double acc = 0.0;
        for (long long it = blockIdx.x; it < ri.items; it += gridDim.x) {
            int len;
            const long long o = ri.locate(w, it, s1, s2, len);
            if (vec) {
#pragma unroll 2
                for (int i = 2 * threadIdx.x; i < len; i += 512) {
                    if (i + 1 < len) {
                        const double2 pv = *reinterpret_cast<const double2*>(p + o + i), qv = *reinterpret_cast<const double2*>(q + o + i);
                        double2 xv = *reinterpret_cast<double2*>(x + o + i), rv = *reinterpret_cast<double2*>(r + o + i);
                        xv.x += alpha * pv.x, xv.y += alpha * pv.y;
                        rv.x -= alpha * qv.x, rv.y -= alpha * qv.y;
                        *reinterpret_cast<double2*>(x + o + i) = xv;
                        *reinterpret_cast<double2*>(r + o + i) = rv;
                        acc += rv.x * rv.x + rv.y * rv.y;
                    } else {
                        x[o + i] += alpha * p[o + i];
                        const double rn = r[o + i] - alpha * q[o + i];
                        r[o + i] = rn;
                        acc += rn * rn;
                    }
                }
            } else {
                for (int i = threadIdx.x; i < len; i += 256) {
                    x[o + i] += alpha * p[o + i];
                    const double rn = r[o + i] - alpha * q[o + i];
                    r[o + i] = rn;
                    acc += rn * rn;
                }
            }
        }
        acc = opf::block_reduce(0, acc);
        if (threadIdx.x == 0) partials[blockIdx.x] = acc;
    }
    // partial sums of a.b over a box (the Krylov dot products; the generic reduction of an expression stays opf::reduce_kernel)
    __global__ void __launch_bounds__(256) dot_fast_kernel(const double* __restrict__ a, const double* __restrict__ b, long long s1, long long s2,
                                                           opf::LaunchRange w, double* __restrict__ partials, int vec) {
        const RowIter ri(w);
        double acc0 = 0.0, acc1 = 0.0;
        for (long long it = blockIdx.x; it < ri.items; it += gridDim.x) {
            int len;
            const long long o = ri.locate(w, it, s1, s2, len);
            if (vec) {
#pragma unroll 2
                for (int i = 2 * threadIdx.x; i < len; i += 512) {
                    if (i + 1 < len) {
                        const double2 av = *reinterpret_cast<const double2*>(a + o + i), bv = *reinterpret_cast<const double2*>(b + o + i);
                        acc0 += av.x * bv.x;
                        acc1 += av.y * bv.y;
                    } else
                        acc0 += a[o + i] * b[o + i];
                }
            } else {
                for (int i = threadIdx.x; i < len; i += 256) acc0 += a[o + i] * b[o + i];
            }
        }
        const double acc = opf::block_reduce(0, acc0 + acc1);
        if (threadIdx.x == 0) partials[blockIdx.x] = acc;
    }
    // fixed-order fold of the partials into *out (deterministic run to run)
    __global__ void __launch_bounds__(256) fold_kernel(const double* __restrict__ partials, int n, double* __restrict__ out) {
        double acc = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partials[i];
        acc = opf::block_reduce(0, acc);
        if (threadIdx.x == 0) out[0] = acc;
    }
    // folds the partials into ks[KS_RR], counts the iteration, decides continuation
    __global__ void __launch_bounds__(256) pcg_control_kernel(const double* __restrict__ partials, int n, double* __restrict__ ks, PcgCtl* ctl,
                                                              cudaGraphConditionalHandle handle, int use_cond, int allreduce_pending) {
        double acc = 0.0;
        if (partials) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partials[i];
            acc = opf::block_reduce(0, acc);
        }
        if (threadIdx.x == 0) {
            if (partials) ks[KS_RR] = acc;
            if (!allreduce_pending) {
                const double rr = ks[KS_RR];
                const bool brk = ks[KS_PQ] == 0.0;// p.q == 0: the update kernel took alpha = 0, nothing moved
                if (!brk) ctl->iters += 1;
                const double rel = sqrt(rr) / ks[KS_BNORM];
                ctl->rel = rel;
                ctl->breakdown = brk;
                const int cont = (rel > ks[KS_TOL]) && (ctl->iters < ctl->maxit) && !brk;
                ctl->cont = cont;
                if (use_cond) cudaGraphSetConditional(handle, cont ? 1u : 0u);
            }
        }
    }
    // p = z + beta p with beta = (r.z)_new / (r.z)_old
    __global__ void __launch_bounds__(256) pcg_direction_kernel(double* __restrict__ p, const double* __restrict__ z, long long s1, long long s2,
                                                                opf::LaunchRange w, const double* __restrict__ ks, int vec) {
        const double beta = ks[KS_RZN] / ks[KS_RZ];
        const RowIter ri(w);
        for (long long it = blockIdx.x; it < ri.items; it += gridDim.x) {
            int len;
            const long long o = ri.locate(w, it, s1, s2, len);
            if (vec) {
#pragma unroll 2
                for (int i = 2 * threadIdx.x; i < len; i += 512) {
                    if (i + 1 < len) {
                        const double2 zv = *reinterpret_cast<const double2*>(z + o + i);
                        double2 pv = *reinterpret_cast<double2*>(p + o + i);
                        pv.x = zv.x + beta * pv.x, pv.y = zv.y + beta * pv.y;
                        *reinterpret_cast<double2*>(p + o + i) = pv;
                    } else
                        p[o + i] = z[o + i] + beta * p[o + i];
                }
            } else {
                for (int i = threadIdx.x; i < len; i += 256) p[o + i] = z[o + i] + beta * p[o + i];
            }
        }
    }
    __global__ void pcg_shift_kernel(double* ks) { ks[KS_RZ] = ks[KS_RZN]; }

    // box launch: threads along axis 0, rows and planes from blockIdx.y / .z
    struct BoxGrid {
        dim3 grid, block;
    };
    BoxGrid box_grid(const Range& w) {
        const int n0 = std::max(1, w.end[0] - w.start[0]), n1 = std::max(1, w.end[1] - w.start[1]), n2 = std::max(1, w.end[2] - w.start[2]);
        int bs = 256;
        while (bs > 32 && bs / 2 >= n0) bs >>= 1;
        return BoxGrid{dim3((n0 + bs - 1) / bs, n1, n2), dim3(bs, 1, 1)};
    }
    opf::LaunchRange lr_of(const Range& r) {
        opf::LaunchRange o;
        for (int d = 0; d < 3; ++d) o.lo[d] = r.start[d], o.hi[d] = r.end[d];
        return o;
    }
}// namespace

struct opf_solver_s {
    struct Level {
        opf_mesh_s* mesh = nullptr;
        opf_field_s *x = nullptr, *b = nullptr, *r = nullptr, *q = nullptr, *dinv = nullptr;
        Range w;            // cells this rank writes on this level
        Range g;            // writable cells of the whole level (== w on one rank and on replicated levels)
        bool dist = false;  // fields are slab/block-decomposed like the target (halo exchange in updatePadding, global reductions)
        bool owns_pin = true;// the first assignable cell lies in w
        long long pin_off = 0;// first assignable cell of this level (pinned on every level when the problem is pinned)
        // transfer tables between this level and the next coarser one (built on first use)
        XferTab tab{};
        void* tab_mem = nullptr;
        bool tab_ready = false;
        // coarse levels of an operator with coefficient fields: one restricted copy per non-unknown lhs leaf (same slot order as
        // lhs_fields; empty on level 0, which reads the caller's fields), and the constant part of an affine lhs on this level
        std::vector<opf_field_s*> coef;
        opf_field_s* c0 = nullptr;
    };
    opf_field_s* target = nullptr;
    std::string lhs_sig, res_sig, smooth_sig, rb_sig[2];
    bool has_smooth_sig = false, has_rb_sig = false;
    std::vector<int> smooth_coef_leaf;
    std::vector<opf_field_s*> lhs_fields;
    std::vector<double> lhs_scalars;
    unsigned mask = 0;
    opf_solver_params params{};
    std::vector<Level> lv;
    opf_field_s *X = nullptr, *B = nullptr, *R = nullptr, *P = nullptr, *Z = nullptr, *Q = nullptr;
    opf_field_s *R0 = nullptr, *V = nullptr, *S = nullptr, *T = nullptr, *E0 = nullptr;// BiCGSTAB extras, boundary-data field
    std::vector<opf_field_s*> gm_v;// GMRES basis (k + 1 vectors, allocated on first use)
    int lattice[4] = {0, 0, 0, 0};   // probe colouring of the diagonal extraction: (a, b, c, M), found once
    bool lattice_failed = false;
    int solves = 0;                  // opf_solver_solve calls so far (lagged set-up of non-static operators)
    // an `lhs` that also carries terms without the unknown (the front-end passes  lhs(e) - rhs(e)  when both sides of `==`
    // contain e) is affine: lhs(p) = A.p + c.  C0 = lhs(0 with homogeneous BCs) = c is removed from every operator application.
    opf_field_s* C0 = nullptr;
    bool affine = false;
    // the multigrid preconditioner M^-1 r -> z is a fixed sequence of ~250 small launches (12 levels at 4097^2) with no host
    // decision inside: after a warm-up call it is captured once per (r, z) pair into a CUDA graph and replayed
    struct VGraph {
        opf_field_s *r, *z;
        unsigned parity;// bit l: which buffer of level l's iterate is current (the fused sweep ping-pongs it)
        unsigned parity_after;// the same after the V-cycle: a replay must flip the host-side `cur` flags like the captured run did
        int mode;             // arithmetic mode the captured kernels were instantiated for (opf_set_mode may change between solves)
        cudaGraphExec_t exec;
        long long launches;
        int calls;
    };
    std::vector<VGraph> vgraphs;
    // device-resident PCG state: scalars, control block (mapped pinned host memory), the captured WHILE-loop graphs
    double* ks = nullptr;
    PcgCtl* ctl_host = nullptr;
    PcgCtl* ctl_dev = nullptr;
    struct LoopGraph {
        unsigned parity, parity_after;
        int mode, precond, pin_active;
        cudaGraphExec_t exec;
        long long launches;
    };
    std::vector<LoopGraph> loops;
    bool in_loop_capture = false;
    bool pinned = false;
    long long pin_off = 0;
    bool setup_done = false, mg = false, has_res_sig = false;
    bool coef_mg = false;   // the hierarchy carries restricted coefficient fields (lhs has leaves besides the unknown)
    bool mg_bypass = false; // coefficient operator whose diagonal dominates (Jacobi is enough): the V-cycle stops at level 0
    bool singular = false;  // pinned AND the operator annihilates constants (all-Neumann / periodic)
    bool pin_active = false;// the Krylov-level operator currently carries the identity row of the pinned unknown
    double omega = 0.8;
};

namespace {
    using Solver = opf_solver_s;

    opf_field_s* clone_homogeneous(opf_field_s* f, const char* name) {
        opf_field_s* c = opf_field_clone(f, name);
        if (!c) return nullptr;
        for (int d = 0; d < c->dim; ++d)
            for (int s = 0; s < 2; ++s) {
                c->bc[d][s].value = 0.0;
                if (c->bc[d][s].face_dev) {
                    cudaFree(c->bc[d][s].face_dev);
                    c->bc[d][s].face_dev = nullptr;
                    c->bc[d][s].face.clear();
                }
            }
        cudaMemsetAsync(c->buf[0], 0, sizeof(double) * c->elems, ctx().stream);
        c->bc0_clean[0] = c->bc0_clean[1] = false;
        return c;
    }

    int assign(opf_field_s* dst, const char* sig, std::initializer_list<opf_field_s*> fs, std::initializer_list<double> sc) {
        opf_field_t F[8];
        double S[8];
        int nf = 0, ns = 0;
        for (auto f : fs) F[nf++] = f;
        for (auto s : sc) S[ns++] = s;
        return opf_assign_ex(dst, OPF_OP_EQ, sig, F, nf, S, ns, OPF_ASSIGN_NO_PADDING);
    }
    // q = lhs(in): ghost fill of `in` with its (homogeneous) BCs, then the expression functor
    // raw: the expression itself, constant part included (used once per solve to build b)
    // which buffer of every caller-owned coefficient field is current (an aliased stencil assignment to such a field flips it):
    // part of the key of every captured graph, whose launches carry the buffer addresses.  Bits 16.. ; the low 16 are the levels'.
    unsigned coef_parity(const Solver* s) {
        unsigned pp = 0;
        for (size_t k = 0; k < s->lhs_fields.size() && k < 16; ++k)
            if (!((s->mask >> k) & 1u) && s->lhs_fields[k]) pp |= (unsigned) (s->lhs_fields[k]->cur & 1) << (16 + k);
        return pp;
    }
    // the coefficient leaf k of lhs as level `level` sees it
    opf_field_s* leaf(Solver* s, int level, int k) { return (level > 0 && !s->lv[level].coef.empty()) ? s->lv[level].coef[k] : s->lhs_fields[k]; }
    int apply_lhs(Solver* s, opf_field_s* in, opf_field_s* out, int level, bool pin = true, bool raw = false) {
        if (int rc = field_update_padding(in)) return rc;
        opf_field_t F[OPF_MAX_FIELDS];
        const int nf = (int) s->lhs_fields.size();
        for (int k = 0; k < nf; ++k) F[k] = ((s->mask >> k) & 1u) ? in : leaf(s, level, k);
        if (int rc = opf_assign_ex(out, OPF_OP_EQ, s->lhs_sig.c_str(), F, nf, s->lhs_scalars.data(), (int) s->lhs_scalars.size(), OPF_ASSIGN_NO_PADDING))
            return rc;
        if (s->affine && !raw) {
            opf_field_s* c0 = level == 0 ? s->C0 : s->lv[level].c0;
            if (c0)
                if (int rc = assign(out, "Sub<F<0>,F<1>>", {out, c0}, {})) return rc;
        }
        if (s->pin_active && pin && s->lv[level].owns_pin) {// identity row for the pinned unknown (HYPREEqnSolveHandler.hpp:145-163)
            copy_cell_kernel<<<1, 1, 0, ctx().stream>>>(out->biased(out->cur), in->biased(in->cur), s->lv[level].pin_off);
            ctx().launches++;
        }
        return OPF_OK;
    }
    // r = b - lhs(x)
    int residual(Solver* s, opf_field_s* x, opf_field_s* b, opf_field_s* r, opf_field_s* scratch, int level, bool pin = true) {
        if (s->has_res_sig && !(s->pin_active && pin)) {
            // affine lhs = A + c:  b - A x = (b + c) - lhs(x) -- the shifted right-hand side goes through the scratch field
            opf_field_s* c0 = s->affine ? (level == 0 ? s->C0 : s->lv[level].c0) : nullptr;
            if (c0) {
                if (int rc = assign(scratch, "Add<F<0>,F<1>>", {b, c0}, {})) return rc;
                b = scratch;
            }
            if (int rc = field_update_padding(x)) return rc;
            opf_field_t F[OPF_MAX_FIELDS];
            const int nf = (int) s->lhs_fields.size();
            F[0] = b;
            for (int k = 0; k < nf; ++k) F[k + 1] = ((s->mask >> k) & 1u) ? x : leaf(s, level, k);
            return opf_assign_ex(r, OPF_OP_EQ, s->res_sig.c_str(), F, nf + 1, s->lhs_scalars.data(), (int) s->lhs_scalars.size(), OPF_ASSIGN_NO_PADDING);
        }
        if (int rc = apply_lhs(s, x, scratch, level, pin)) return rc;
        return assign(r, "Sub<F<0>,F<1>>", {b, scratch}, {});
    }
    int dot(Solver* s, opf_field_s* a, opf_field_s* b, const Range& w, double* out) {
        opf_field_t F[2] = {a, b};
        opf_range r = to_c(w);
        if (int rc = opf_reduce(OPF_RED_SUM, "Mul<F<0>,F<1>>", F, 2, nullptr, 0, &r, out)) return rc;
        if (s->target->n_ranks > 1 && comm_active()) return opf_comm_allreduce(out, 1, OPF_RED_SUM);
        return OPF_OK;
    }
    void poke(Solver* s, opf_field_s* f, long long off, double v) {// the pinned cell lives on one rank only
        if (!s->lv[0].owns_pin) return;
        set_cell_kernel<<<1, 1, 0, ctx().stream>>>(f->biased(f->cur), off, v);
        ctx().launches++;
    }

    int build_diag(Solver* s, int level) {
        auto& L = s->lv[level];
        const int dim = s->target->dim;
        const int m = 2 * signature_radius(s->lhs_sig.c_str()) + 1;
        // probe colouring: cells closer than the stencil width must differ in colour, also across a periodic seam
        Mod3 mm{{1, 1, 1}, {0, 0, 0, 0}};
        for (int d = 0; d < dim; ++d) {
            mm.m[d] = m;
            if (L.x->bc[d][0].type == OPF_BC_PERIODIC) {
                const int per = L.x->accessible.end[d] - L.x->accessible.start[d];
                while (per % mm.m[d] != 0 && per % mm.m[d] < m && mm.m[d] < per) mm.m[d]++;
            }
        }
        // Lattice colouring: cells i, j may share a probe iff no row reads both, i.e. i - j is not a difference of two footprint
        // offsets.  colour = (a i + b j + c k) mod M separates them when a dx + b dy + c dz != 0 (mod M) for every such difference;
        // the smallest M found (a brute-force search over M <= 128, once per solver) replaces the (2r+1)^dim cube: 27 -> 7 probes for
        // the 7-point Laplacian, 125 -> ~20 for the semi-implicit momentum operators.  A periodic axis of extent N needs coef * N == 0
        // (mod M) so that the colouring closes across the seam.
        for (int q = 0; q < 4; ++q) mm.lat[q] = 0;
        static const int lattice_on = getenv("OPF_DIAG_LATTICE") ? atoi(getenv("OPF_DIAG_LATTICE")) : 1;
        if (lattice_on) {
            if (s->lattice[3] == 0 && !s->lattice_failed) {
                std::vector<std::array<int, 3>> taps;
                bool found = false;
                if (!signature_unknown_taps(s->lhs_sig.c_str(), s->mask, taps) && !taps.empty() && taps.size() <= 64) {
                    std::vector<std::array<int, 3>> diff;
                    for (const auto& p1 : taps)
                        for (const auto& p2 : taps)
                            if (p1 != p2) diff.push_back({p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]});
                    std::sort(diff.begin(), diff.end());
                    diff.erase(std::unique(diff.begin(), diff.end()), diff.end());
                    long long per0[3] = {0, 0, 0};// periodic extents of the finest level (0: axis not periodic)
                    for (int d = 0; d < dim; ++d)
                        if (s->lv[0].x->bc[d][0].type == OPF_BC_PERIODIC) per0[d] = s->lv[0].x->accessible.end[d] - s->lv[0].x->accessible.start[d];
                    for (int M = (int) taps.size(); M <= 96 && !found; ++M)
                        for (int a = 1; a < std::min(std::max(2, M), 9) && !found; ++a)// a small first coefficient is enough in practice
                            for (int b = (dim >= 2 ? 1 : 0); b < (dim >= 2 ? M : 1) && !found; ++b)
                                for (int c = (dim >= 3 ? 1 : 0); c < (dim >= 3 ? M : 1) && !found; ++c) {
                                    if ((a * per0[0]) % M || (b * per0[1]) % M || (c * per0[2]) % M) continue;
                                    bool ok = true;
                                    for (const auto& d : diff) {
                                        long long v = ((long long) a * d[0] + (long long) b * d[1] + (long long) c * d[2]) % M;
                                        if (v == 0) {
                                            ok = false;
                                            break;
                                        }
                                    }
                                    if (ok) {
                                        s->lattice[0] = a, s->lattice[1] = b, s->lattice[2] = c, s->lattice[3] = M;
                                        found = true;
                                    }
                                }
                }
                if (!found) s->lattice_failed = true;
            }
            if (s->lattice[3] > 0) {
                bool closes = true;// every level: periodic extents must be compatible with the lattice
                const int coef[3] = {s->lattice[0], s->lattice[1], s->lattice[2]};
                for (int d = 0; d < dim; ++d)
                    if (L.x->bc[d][0].type == OPF_BC_PERIODIC) {
                        const long long per = L.x->accessible.end[d] - L.x->accessible.start[d];
                        if ((coef[d] * per) % s->lattice[3] != 0) closes = false;
                    }
                if (closes)
                    for (int q = 0; q < 4; ++q) mm.lat[q] = s->lattice[q];
            }
        }
        const opf::LaunchRange r = lr_of(L.w);
        const BoxGrid bg = box_grid(L.w);
        if (mm.lat[3] > 0) {
            for (int col = 0; col < mm.lat[3]; ++col) {
                color_fill_kernel<<<bg.grid, bg.block, 0, ctx().stream>>>(L.x->biased(L.x->cur), L.x->pitch1, L.x->pitch2, r, dim, mm, col, 0, 0);
                ctx().launches++;
                if (int rc = apply_lhs(s, L.x, L.q, level, false)) return rc;
                color_recip_kernel<<<bg.grid, bg.block, 0, ctx().stream>>>(L.q->biased(L.q->cur), L.dinv->biased(L.dinv->cur), L.dinv->pitch1, L.dinv->pitch2, r, dim, mm,
                                                                           col, 0, 0);
                ctx().launches++;
            }
            OPF_CUDA(cudaGetLastError());
            return assign(L.x, "S<0>", {}, {0.0});
        }
        for (int c2 = 0; c2 < mm.m[2]; ++c2)
            for (int c1 = 0; c1 < mm.m[1]; ++c1)
                for (int c0 = 0; c0 < mm.m[0]; ++c0) {
                    color_fill_kernel<<<bg.grid, bg.block, 0, ctx().stream>>>(L.x->biased(L.x->cur), L.x->pitch1, L.x->pitch2, r, dim, mm, c0, c1, c2);
                    ctx().launches++;
                    if (int rc = apply_lhs(s, L.x, L.q, level, false)) return rc;
                    color_recip_kernel<<<bg.grid, bg.block, 0, ctx().stream>>>(L.q->biased(L.q->cur), L.dinv->biased(L.dinv->cur), L.dinv->pitch1, L.dinv->pitch2, r,
                                                                     dim, mm, c0, c1, c2);
                    ctx().launches++;
                }
        OPF_CUDA(cudaGetLastError());
        return assign(L.x, "S<0>", {}, {0.0});
    }

    // relax_type 0 / 1: (weighted) Jacobi.  2 / 3: red-black Gauss-Seidel (StructSolverPFMG.hpp:23-34) -- two half-sweeps with the exact
    // diagonal on the cells of one (i + j + k) parity each; 2 is the symmetric variant (red-black before, black-red after the coarse
    // correction, which keeps the V-cycle a symmetric preconditioner for PCG), 3 runs red-black on both sides.
    int smooth(Solver* s, int level, int sweeps, bool zero_guess, bool post = false) {
        auto& L = s->lv[level];
        const int rt = s->params.relax_type;
        if (rt == 2 || rt == 3) {
            const int first = (post && rt == 2) ? 1 : 0;
            for (int it = 0; it < sweeps; ++it)
                for (int h = 0; h < 2; ++h) {
                    const int c = h == 0 ? first : 1 - first;
                    if (it == 0 && h == 0 && zero_guess && c == 0) {// x = 0: the red cells get dinv * b, the black ones stay 0
                        if (int rc = assign(L.x, "Mul<Par<0>,Mul<F<0>,F<1>>>", {L.dinv, L.b}, {})) return rc;
                        continue;
                    }
                    if (it == 0 && h == 0 && zero_guess)
                        if (int rc = assign(L.x, "S<0>", {}, {0.0})) return rc;
                    if (s->has_rb_sig && !s->affine) {
                        if (int rc = field_update_padding(L.x)) return rc;
                        opf_field_t F[OPF_MAX_FIELDS];
                        double S[OPF_MAX_SCALARS];
                        const int nc = (int) s->smooth_coef_leaf.size(), ns = (int) s->lhs_scalars.size();
                        F[0] = L.x, F[1] = L.dinv, F[2] = L.b;
                        for (int k = 0; k < nc; ++k) F[k + 3] = leaf(s, level, s->smooth_coef_leaf[k]);
                        S[0] = 1.0;
                        for (int k = 0; k < ns; ++k) S[k + 1] = s->lhs_scalars[k];
                        if (int rc = opf_assign_ex(L.x, OPF_OP_EQ, s->rb_sig[c].c_str(), F, nc + 3, S, ns + 1, OPF_ASSIGN_NO_PADDING)) return rc;
                    } else {
                        if (int rc = residual(s, L.x, L.b, L.r, L.q, level, false)) return rc;
                        if (int rc = assign(L.x, c == 0 ? "Add<F<0>,Mul<Par<0>,Mul<F<1>,F<2>>>>" : "Add<F<0>,Mul<Par<1>,Mul<F<1>,F<2>>>>", {L.x, L.dinv, L.r}, {})) return rc;
                    }
                }
            return OPF_OK;
        }
        for (int it = 0; it < sweeps; ++it) {
            if (it == 0 && zero_guess) {
                if (int rc = assign(L.x, "Mul<S<0>,Mul<F<0>,F<1>>>", {L.dinv, L.b}, {s->omega})) return rc;
            } else if (s->has_smooth_sig && !s->affine) {
                // one sweep, one pass: reads x (stencil), dinv, b; writes the twin of x (ping-pong) -- 32 B per cell instead of 56
                if (int rc = field_update_padding(L.x)) return rc;
                opf_field_t F[OPF_MAX_FIELDS];
                double S[OPF_MAX_SCALARS];
                const int nc = (int) s->smooth_coef_leaf.size(), ns = (int) s->lhs_scalars.size();
                F[0] = L.x, F[1] = L.dinv, F[2] = L.b;
                for (int k = 0; k < nc; ++k) F[k + 3] = leaf(s, level, s->smooth_coef_leaf[k]);
                S[0] = s->omega;
                for (int k = 0; k < ns; ++k) S[k + 1] = s->lhs_scalars[k];
                if (int rc = opf_assign_ex(L.x, OPF_OP_EQ, s->smooth_sig.c_str(), F, nc + 3, S, ns + 1, OPF_ASSIGN_NO_PADDING)) return rc;
            } else {
                if (int rc = residual(s, L.x, L.b, L.r, L.q, level, false)) return rc;
                if (int rc = assign(L.x, "Add<F<0>,Mul<S<0>,Mul<F<1>,F<2>>>>", {L.x, L.dinv, L.r}, {s->omega})) return rc;
            }
        }
        return OPF_OK;
    }

    XferParams xfer(Solver* s, int lf) {
        auto &Lf = s->lv[lf], &Lc = s->lv[lf + 1];
        XferParams p{};
        p.dim = s->target->dim;
        for (int d = 0; d < 3; ++d) {
            p.flo[d] = Lf.g.start[d], p.fhi[d] = Lf.g.end[d];
            p.clo[d] = Lc.g.start[d], p.chi[d] = Lc.g.end[d];
            p.wlo[d] = 0, p.whi[d] = 1;
            const bool per = d < p.dim && Lf.x->bc[d][0].type == OPF_BC_PERIODIC;
            p.fhalo[d] = per && Lf.dist;
            p.chalo[d] = per && Lc.dist;
            p.f0[d] = Lf.x->accessible.start[d], p.c0[d] = Lc.x->accessible.start[d];
            p.center[d] = d < p.dim ? (Lf.x->loc[d] == OPF_LOC_CENTER || Lf.x->bc[d][0].type == OPF_BC_PERIODIC) : 0;
            p.periodic[d] = d < p.dim && Lf.x->bc[d][0].type == OPF_BC_PERIODIC;
            p.bclo[d] = d < p.dim ? Lf.x->bc[d][0].type : 0;
            p.bchi[d] = d < p.dim ? Lf.x->bc[d][1].type : 0;
            p.fper[d] = Lf.x->accessible.end[d] - Lf.x->accessible.start[d];
            p.cper[d] = Lc.x->accessible.end[d] - Lc.x->accessible.start[d];
        }
        // periodic Corner fields are node-based with period n-1: their 2:1 map is the vertex one
        for (int d = 0; d < p.dim; ++d)
            if (p.periodic[d] && Lf.x->loc[d] == OPF_LOC_CORNER) p.center[d] = 0;
        return p;
    }


    // host evaluation of the per-axis transfer stencils into device tables covering the whole storage range of both levels
    int build_xfer_tab(Solver* s, int lf) {
        auto &Lf = s->lv[lf], &Lc = s->lv[lf + 1];
        if (Lf.tab_ready) return OPF_OK;
        const XferParams p = xfer(s, lf);
        std::vector<int> ib;
        std::vector<double> wb;
        size_t ioff[3][2], woff[3][2];
        for (int d = 0; d < 3; ++d) {
            AxisTab& a = Lf.tab.ax[d];
            const bool used = d < p.dim;
            a.plo = used ? Lf.x->storage.start[d] : 0;
            a.pn = used ? Lf.x->storage.end[d] - Lf.x->storage.start[d] : 1;
            a.rlo = used ? Lc.x->storage.start[d] : 0;
            a.rn = used ? Lc.x->storage.end[d] - Lc.x->storage.start[d] : 1;
            a.rcnt = used ? (p.center[d] ? 4 : 3) : 1;
            ioff[d][0] = ib.size(), woff[d][0] = wb.size();
            for (int q = 0; q < a.pn; ++q) {
                int idx[2];
                double w[2];
                host_prolong_axis(p, d, a.plo + q, idx, w);
                ib.push_back(idx[0]), ib.push_back(idx[1]), wb.push_back(w[0]), wb.push_back(w[1]);
            }
            ioff[d][1] = ib.size(), woff[d][1] = wb.size();
            for (int q = 0; q < a.rn; ++q) {
                int idx[4];
                double w[4];
                host_restrict_axis(p, d, a.rlo + q, idx, w);
                for (int t = 0; t < 4; ++t) ib.push_back(idx[t]), wb.push_back(w[t]);
            }
        }
        const size_t wbytes = wb.size() * sizeof(double), ibytes = ib.size() * sizeof(int);
        OPF_CUDA(cudaMalloc(&Lf.tab_mem, wbytes + ibytes));
        OPF_CUDA(cudaMemcpy(Lf.tab_mem, wb.data(), wbytes, cudaMemcpyHostToDevice));
        OPF_CUDA(cudaMemcpy((char*) Lf.tab_mem + wbytes, ib.data(), ibytes, cudaMemcpyHostToDevice));
        const double* wd = (const double*) Lf.tab_mem;
        const int* id = (const int*) ((char*) Lf.tab_mem + wbytes);
        for (int d = 0; d < 3; ++d) {
            AxisTab& a = Lf.tab.ax[d];
            a.pidx = id + ioff[d][0], a.pw = wd + woff[d][0];
            a.ridx = id + ioff[d][1], a.rw = wd + woff[d][1];
        }
        Lf.tab_ready = true;
        return OPF_OK;
    }

    // singular operators (the caller pinned a value: all-Neumann / periodic): keep every level's right-hand side in the
    // range of the operator by removing its mean -- on the device, no host round trip
    int project_mean(Solver* s, opf_field_s* f, const Range& w, const Range& g, bool dist) {
        double* dev = nullptr;
        if (int rc = reduce_sum_device(f, w, &dev)) return rc;
        if (dist)
            if (int rc = comm_allreduce_device(dev, 1, OPF_RED_SUM, ctx().stream)) return rc;
        sub_mean_kernel<<<box_grid(w).grid, box_grid(w).block, 0, ctx().stream>>>(f->biased(f->cur), f->pitch1, f->pitch2, lr_of(w), dev, 1.0 / (double) g.count());
        ctx().launches++;
        return OPF_OK;
    }

    // top_is_mean_free: the caller guarantees a mean-free right-hand side on this level (the Krylov residual of the singular
    // phase) -- its projection pass is skipped
    int vcycle(Solver* s, int level, bool zero_guess, bool top_is_mean_free = false) {
        auto& L = s->lv[level];
        const int last = (int) s->lv.size() - 1;
        // Coarse right-hand sides are restrictions of a mean-free residual by an operator whose columns all sum to 2^-dim (R = P^T / 2^dim,
        // P interpolates constants exactly on all-Neumann / periodic problems): mean-free again up to rounding, and nothing accumulates
        // because every cycle starts from the Krylov residual.  Their projection passes (3 launches per level per cycle) are off by
        // default; OPF_MG_PROJECT_COARSE=1 restores them.
        static const int project_coarse = getenv("OPF_MG_PROJECT_COARSE") ? atoi(getenv("OPF_MG_PROJECT_COARSE")) : 0;
        if (s->singular && !top_is_mean_free && (level == 0 || project_coarse))
            if (int rc = project_mean(s, L.b, L.w, L.g, L.dist)) return rc;
        static const int coarse_sweeps = getenv("OPF_MG_COARSE_SWEEPS") ? atoi(getenv("OPF_MG_COARSE_SWEEPS")) : 8;
        if (s->mg_bypass && level == 0) return smooth(s, 0, std::max(1, s->params.num_pre_relax), zero_guess);
        if (level == last) return smooth(s, level, last == 0 ? std::max(1, s->params.num_pre_relax) : coarse_sweeps, zero_guess);
        const int pre = std::max(1, s->params.num_pre_relax), post = std::max(1, s->params.num_post_relax);
        if (int rc = smooth(s, level, pre, zero_guess)) return rc;
        if (int rc = residual(s, L.x, L.b, L.r, L.q, level, false)) return rc;
        auto& C = s->lv[level + 1];
        XferParams p = xfer(s, level);
        // ---- restriction.  Decomposed fine level: the stencil reaches one cell into the neighbours' blocks -> refresh r's halo.
        if (L.dist)
            if (int rc = field_update_padding(L.r)) return rc;
        p.fine = L.r->biased(L.r->cur), p.fs1 = L.r->pitch1, p.fs2 = L.r->pitch2;
        p.coarse = C.b->biased(C.b->cur), p.cs1 = C.b->pitch1, p.cs2 = C.b->pitch2;
        Range cw = C.w;
        if (L.dist && !C.dist) {
            // transition to the replicated coarse hierarchy: every rank restricts the coarse cells under its own fine block into a
            // zeroed full-size coarse vector, one allreduce makes it whole on every rank
            for (int d = 0; d < 3; ++d) {
                cw.start[d] = std::max(C.g.start[d], C.x->accessible.start[d] + (L.w.start[d] - L.x->accessible.start[d] + 1) / 2);
                cw.end[d] = std::min(C.g.end[d], C.x->accessible.start[d] + (L.w.end[d] - L.x->accessible.start[d] + 1) / 2);
            }
            OPF_CUDA(cudaMemsetAsync(C.b->buf[C.b->cur], 0, sizeof(double) * C.b->elems, ctx().stream));
        }
        for (int d = 0; d < 3; ++d) p.wlo[d] = cw.start[d], p.whi[d] = cw.end[d];
        static const int fast_xfer = getenv("OPF_MG_FAST_XFER") ? atoi(getenv("OPF_MG_FAST_XFER")) : 1;
        const bool fast = fast_xfer && p.dim >= 2;
        if (fast)
            if (int rc = build_xfer_tab(s, level)) return rc;
        if (cw.count() > 0) {
            if (fast && p.dim == 2) restrict_fast_kernel<2><<<box_grid(cw).grid, box_grid(cw).block, 0, ctx().stream>>>(p.fine, p.fs1, p.fs2, p.coarse, p.cs1, p.cs2, lr_of(cw), L.tab);
            else if (fast)
                restrict_fast_kernel<3><<<box_grid(cw).grid, box_grid(cw).block, 0, ctx().stream>>>(p.fine, p.fs1, p.fs2, p.coarse, p.cs1, p.cs2, lr_of(cw), L.tab);
            else
                restrict_kernel<<<box_grid(cw).grid, box_grid(cw).block, 0, ctx().stream>>>(p);
            ctx().launches++;
        }
        if (L.dist && !C.dist)
            if (int rc = comm_allreduce_device(C.b->buf[C.b->cur], (int) C.b->elems, OPF_RED_SUM, ctx().stream)) return rc;
        if (int rc = vcycle(s, level + 1, true)) return rc;
        // ---- prolongation: a fine cell interpolates from its parent and the parent's neighbour -> coarse halo on decomposed levels
        if (C.dist)
            if (int rc = field_update_padding(C.x)) return rc;
        p.fine = L.x->biased(L.x->cur), p.fs1 = L.x->pitch1, p.fs2 = L.x->pitch2;
        p.coarse = C.x->biased(C.x->cur), p.cs1 = C.x->pitch1, p.cs2 = C.x->pitch2;
        for (int d = 0; d < 3; ++d) p.wlo[d] = L.w.start[d], p.whi[d] = L.w.end[d];
        if (fast && p.dim == 2) prolong_fast_kernel<2><<<box_grid(L.w).grid, box_grid(L.w).block, 0, ctx().stream>>>(p.fine, p.fs1, p.fs2, p.coarse, p.cs1, p.cs2, lr_of(L.w), L.tab);
        else if (fast)
            prolong_fast_kernel<3><<<box_grid(L.w).grid, box_grid(L.w).block, 0, ctx().stream>>>(p.fine, p.fs1, p.fs2, p.coarse, p.cs1, p.cs2, lr_of(L.w), L.tab);
        else
            prolong_kernel<<<box_grid(L.w).grid, box_grid(L.w).block, 0, ctx().stream>>>(p);
        ctx().launches++;
        OPF_CUDA(cudaGetLastError());
        return smooth(s, level, post, false, true);
    }

    int project_mean(Solver* s, opf_field_s* f, const Range& w, const Range& g, bool dist);
    // z = M^-1 r
    int precondition(Solver* s, opf_field_s* r, opf_field_s* z) {
        auto& L0 = s->lv[0];
        switch (s->params.precond) {
            case OPF_SOLVER_PFMG:
            case OPF_SOLVER_SMG: {
                auto body = [&]() -> int {
                    // the V-cycle runs directly on the Krylov vectors: r is its level-0 right-hand side (read only: in the singular
                    // phase it is mean-free already, so the level-0 projection is skipped), z its level-0 iterate -- no copies
                    opf_field_s *sb = L0.b, *sx = L0.x;
                    L0.b = r, L0.x = z;
                    const int cycles = std::max(1, s->params.precond_max_iter);
                    int rc = OPF_OK;
                    for (int c = 0; c < cycles && !rc; ++c) rc = vcycle(s, 0, c == 0, s->singular && !s->pin_active);
                    L0.b = sb, L0.x = sx;
                    return rc;
                };
                const int graphs_on = opf_internal_opt(OPF_OPT_GRAPHS);
                if (!graphs_on || s->lv[0].dist || s->in_loop_capture) return body();// NCCL exchanges inside: not captured; inside the loop capture: inlined
                unsigned parity = (unsigned) z->cur | coef_parity(s);
                for (size_t lv = 1; lv < s->lv.size(); ++lv) parity |= (unsigned) s->lv[lv].x->cur << lv;
                Solver::VGraph* g = nullptr;
                for (auto& e : s->vgraphs)
                    if (e.r == r && e.z == z && e.parity == parity && e.mode == ctx().mode) g = &e;
                if (!g) {
                    s->vgraphs.push_back(Solver::VGraph{r, z, parity, parity, ctx().mode, nullptr, 0, 0});
                    g = &s->vgraphs.back();
                }
                Context& c = ctx();
                auto current_parity = [&]() {
                    unsigned pp = (unsigned) z->cur | coef_parity(s);
                    for (size_t lv = 1; lv < s->lv.size(); ++lv) pp |= (unsigned) s->lv[lv].x->cur << lv;
                    return pp;
                };
                auto apply_parity = [&](unsigned pp) {
                    z->cur = (int) (pp & 1u);
                    for (size_t lv = 1; lv < s->lv.size(); ++lv) s->lv[lv].x->cur = (int) ((pp >> lv) & 1u);
                };
                if (g->exec) {
                    OPF_CUDA(cudaGraphLaunch(g->exec, c.stream));
                    c.launches += g->launches;
                    apply_parity(g->parity_after);
                    return OPF_OK;
                }
                if (++g->calls < 2) return body();// first call: allocations, kernel attribute set-up, BC-clean flags settle
                const long long l0 = c.launches;
                OPF_CUDA(cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeThreadLocal));
                const int rc = body();
                cudaGraph_t graph = nullptr;
                const cudaError_t ce = cudaStreamEndCapture(c.stream, &graph);
                if (rc) {
                    if (graph) cudaGraphDestroy(graph);
                    return rc;
                }
                if (ce != cudaSuccess || !graph) return fail(OPF_ERR_CUDA, "V-cycle graph capture failed: %s", cudaGetErrorString(ce));
                g->launches = c.launches - l0;
                c.launches = l0;
                g->parity_after = current_parity();
                apply_parity(g->parity);// the capture recorded the launches without running them: state is still the entry state
                const cudaError_t ie = cudaGraphInstantiate(&g->exec, graph, 0);
                cudaGraphDestroy(graph);
                if (ie != cudaSuccess) {
                    g->exec = nullptr;
                    return fail(OPF_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ie));
                }
                OPF_CUDA(cudaGraphLaunch(g->exec, c.stream));
                c.launches += g->launches;
                apply_parity(g->parity_after);
                return OPF_OK;
            }
            case OPF_SOLVER_JACOBI: return assign(z, "Mul<F<0>,F<1>>", {L0.dinv, r}, {});
            default: return assign(z, "F<0>", {r}, {});
        }
    }
    int precondition_pinned(Solver* s, opf_field_s* r, opf_field_s* z) {
        if (int rc = precondition(s, r, z)) return rc;
        if (s->pin_active) poke(s, z, s->pin_off, 0.0);// keep the pinned unknown out of the Krylov space (projection P M P)
        else if (s->singular) {
            // singular phase: a constant component of z is invisible to the iteration (r and A.p are mean-free, A annihilates
            // constants) and is removed from x by the pin shift at the end -- the projection pass is only kept as an option
            static const int project_z = getenv("OPF_MG_PROJECT_Z") ? atoi(getenv("OPF_MG_PROJECT_Z")) : 0;
            if (project_z) return project_mean(s, z, s->lv[0].w, s->lv[0].g, s->lv[0].dist);
        }
        return OPF_OK;
    }

    opf_field_s* make_level_field(opf_field_s* like, opf_mesh_s* mesh, const char* name, const std::vector<opf_range>* split = nullptr) {
        opf_field_desc d{};
        d.mesh = mesh;
        for (int a = 0; a < like->dim; ++a) {
            d.loc[a] = like->loc[a];
            for (int sd = 0; sd < 2; ++sd) {
                d.bc[a][sd].type = like->bc[a][sd].type;
                d.bc[a][sd].value = 0.0;
                d.bc[a][sd].face = nullptr;
                d.ext[a][sd] = like->ext[a][sd];
            }
        }
        d.padding = like->padding;
        d.n_ranks = 0;
        if (split) {// same decomposition as the target, block boundaries halved
            d.n_ranks = like->n_ranks;
            d.rank = like->rank;
            d.split_map = split->data();
        }
        return opf_field_create(&d, name);
    }

    void free_level_fields(Solver::Level& L);
    // coarse copy of a coefficient field: same staggering and boundary conditions (real values: the lid speed stays the lid speed)
    opf_field_s* make_coef_field(opf_field_s* like, opf_mesh_s* mesh, const char* name) {
        opf_field_desc d{};
        d.mesh = mesh;
        for (int a = 0; a < like->dim; ++a) {
            d.loc[a] = like->loc[a];
            for (int sd = 0; sd < 2; ++sd) {
                d.bc[a][sd].type = like->bc[a][sd].type;
                d.bc[a][sd].value = like->bc[a][sd].value;
                d.bc[a][sd].face = nullptr;
                d.ext[a][sd] = like->ext[a][sd];
            }
        }
        d.padding = like->padding;
        d.n_ranks = 0;
        return opf_field_create(&d, name);
    }
    // can the hierarchy carry this coefficient field?  same mesh as the unknown, one rank, constant boundary values
    bool coef_coarsenable(const opf_field_s* f, const opf_field_s* t) {
        if (!f || f->dim != t->dim || f->n_ranks > 1) return false;
        for (int d = 0; d < t->dim; ++d) {
            if (f->mesh->dims[d] != t->mesh->dims[d] || f->mesh->range.start[d] != t->mesh->range.start[d]) return false;
            for (int sd = 0; sd < 2; ++sd)
                if (f->bc[d][sd].face_dev) return false;
        }
        return true;
    }

    int build_levels(Solver* s) {
        opf_field_s* t = s->target;
        const int dim = t->dim;
        s->lv.clear();
        Solver::Level L0;
        L0.mesh = t->mesh;
        L0.w = common(t->assignable, t->local);
        L0.g = t->assignable;
        L0.dist = t->n_ranks > 1;
        auto owns = [](const opf_field_s* f, const Range& w) {
            for (int d = 0; d < f->dim; ++d)
                if (f->assignable.start[d] < w.start[d] || f->assignable.start[d] >= w.end[d]) return false;
            return true;
        };
        L0.owns_pin = owns(t, L0.w);
        L0.pin_off = (long long) t->assignable.start[0] + (long long) t->assignable.start[1] * t->pitch1 + (long long) t->assignable.start[2] * t->pitch2;
        L0.x = clone_homogeneous(t, "mg.x0");
        L0.b = clone_homogeneous(t, "mg.b0");
        L0.r = clone_homogeneous(t, "mg.r0");
        L0.q = clone_homogeneous(t, "mg.q0");
        L0.dinv = clone_homogeneous(t, "mg.dinv0");
        if (!L0.x || !L0.b || !L0.r || !L0.q || !L0.dinv) return OPF_ERR_CUDA;
        s->lv.push_back(L0);
        const bool want_mg = s->params.precond == OPF_SOLVER_PFMG || s->params.precond == OPF_SOLVER_SMG || s->params.type == OPF_SOLVER_PFMG
                             || s->params.type == OPF_SOLVER_SMG;
        // multigrid needs an operator that only involves the unknown (coefficient fields are not restricted yet); a decomposed
        // target additionally needs the communicator
        bool pure = true;
        for (size_t k = 0; k < s->lhs_fields.size(); ++k)
            if (!((s->mask >> k) & 1u)) pure = false;
        // ... or whose coefficient fields can be carried down the hierarchy (OPF_MG_COEF: 0 never -- such a request degrades to Jacobi
        // as before --, 1 (default) when the level-0 diagonal does not dominate, 2 always)
        static const int coef_opt = getenv("OPF_MG_COEF") ? atoi(getenv("OPF_MG_COEF")) : 1;
        bool carry = !pure && coef_opt > 0 && t->n_ranks <= 1;
        if (carry)
            for (size_t k = 0; k < s->lhs_fields.size(); ++k)
                if (!((s->mask >> k) & 1u) && !coef_coarsenable(s->lhs_fields[k], t)) carry = false;
        s->coef_mg = want_mg && carry;
        s->mg = want_mg && (pure || carry) && (t->n_ranks <= 1 || (comm_active() && !t->cell_split.empty()));
        if (!s->mg) return OPF_OK;
        // cell-centred blocks of the current distributed level (empty once the hierarchy continues replicated on every rank)
        std::vector<Range> blocks = t->n_ranks > 1 ? t->cell_split : std::vector<Range>();
        for (;;) {
            auto& F = s->lv.back();
            opf_mesh_s* fm = F.mesh;
            bool ok = true;
            int cd[3] = {1, 1, 1};
            for (int d = 0; d < dim; ++d) {
                const int n = fm->dims[d];
                if ((n - 1) % 2 != 0 || (n - 1) / 2 < 2) ok = false;
                cd[d] = (n - 1) / 2 + 1;
            }
            if (!ok || (int) s->lv.size() >= 16) break;
            // can the next level stay decomposed?  every block boundary must be even and every coarse block thick enough for a halo
            std::vector<opf_range> csplit;
            bool cdist = !blocks.empty();
            if (cdist) {
                const int minw = std::max(2, 2 * t->padding);
                for (const auto& b : blocks) {
                    Range c = b;
                    for (int d = 0; d < dim; ++d) {
                        const int rs = b.start[d] - fm->range.start[d], re = b.end[d] - fm->range.start[d];
                        if ((rs & 1) || (re & 1)) cdist = false;
                        c.start[d] = fm->range.start[d] + rs / 2;
                        c.end[d] = fm->range.start[d] + re / 2;
                        const bool split_axis = !(b.start[d] == fm->range.start[d] && b.end[d] == fm->range.end[d] - 1);
                        if (split_axis && c.end[d] - c.start[d] < minw) cdist = false;
                    }
                    csplit.push_back(to_c(c));
                }
            }
            opf_mesh_s* cm = opf_mesh_create(dim, cd, fm->start, fm->pad_width);
            for (int d = 0; d < dim; ++d) {
                cm->ext_mode[d] = fm->ext_mode[d];
                std::vector<double> xs(cd[d]);
                const int off = fm->range.start[d] - fm->ext_range.start[d];
                for (int i = 0; i < cd[d]; ++i) xs[i] = fm->ax[d].x[off + 2 * i];
                if (int rc = opf_mesh_set_coords(cm, d, xs.data(), cd[d])) return rc;
            }
            const std::vector<opf_range>* sp = cdist ? &csplit : nullptr;
            Solver::Level C;
            C.mesh = cm;
            C.dist = cdist;
            C.x = make_level_field(t, cm, "mg.x", sp);
            C.b = make_level_field(t, cm, "mg.b", sp);
            C.r = make_level_field(t, cm, "mg.r", sp);
            C.q = make_level_field(t, cm, "mg.q", sp);
            C.dinv = make_level_field(t, cm, "mg.dinv", sp);
            if (!C.x || !C.b || !C.r || !C.q || !C.dinv) return OPF_ERR_CUDA;
            if (s->coef_mg) {
                C.coef.assign(s->lhs_fields.size(), nullptr);
                for (size_t k = 0; k < s->lhs_fields.size(); ++k) {
                    if ((s->mask >> k) & 1u) continue;
                    for (size_t q = 0; q < k && !C.coef[k]; ++q)// the same field in two slots shares its coarse copy
                        if (!((s->mask >> q) & 1u) && s->lhs_fields[q] == s->lhs_fields[k]) C.coef[k] = C.coef[q];
                    if (!C.coef[k] && !(C.coef[k] = make_coef_field(s->lhs_fields[k], cm, "mg.coef"))) return OPF_ERR_CUDA;
                }
                if (!(C.c0 = make_level_field(t, cm, "mg.c0", sp))) return OPF_ERR_CUDA;
            }
            opf_mesh_destroy(cm);// fields hold their own references
            C.w = common(C.x->assignable, C.x->local);
            C.g = C.x->assignable;
            C.owns_pin = owns(C.x, C.w);
            C.pin_off = (long long) C.x->assignable.start[0] + (long long) C.x->assignable.start[1] * C.x->pitch1
                        + (long long) C.x->assignable.start[2] * C.x->pitch2;
            if (C.g.count() <= 0) {
                free_level_fields(C);
                break;
            }
            s->lv.push_back(C);
            if (cdist) {
                blocks.clear();
                for (const auto& c : csplit) blocks.push_back(from_c(c, dim));
            } else
                blocks.clear();
        }
        return OPF_OK;
    }

    void free_level_fields(Solver::Level& L) {
        if (L.tab_mem) cudaFree(L.tab_mem);
        L.tab_mem = nullptr;
        L.tab_ready = false;
        for (opf_field_s* f : {L.x, L.b, L.r, L.q, L.dinv, L.c0})
            if (f) opf_field_destroy(f);
        L.x = L.b = L.r = L.q = L.dinv = L.c0 = nullptr;
        for (size_t k = 0; k < L.coef.size(); ++k) {
            if (!L.coef[k]) continue;
            opf_field_destroy(L.coef[k]);
            for (size_t q = k + 1; q < L.coef.size(); ++q)
                if (L.coef[q] == L.coef[k]) L.coef[q] = nullptr;
            L.coef[k] = nullptr;
        }
    }

    // refresh the coarse copies of the coefficient fields (level l from level l - 1), their ghosts, and the constant part of an
    // affine lhs on every coarse level
    int restrict_coefs(Solver* s) {
        const int dim = s->target->dim;
        for (int l = 1; l < (int) s->lv.size(); ++l) {
            auto &F = s->lv[l - 1], &C = s->lv[l];
            for (size_t k = 0; k < C.coef.size(); ++k) {
                opf_field_s* dst = C.coef[k];
                if (!dst) continue;
                bool seen = false;
                for (size_t q = 0; q < k; ++q) seen = seen || C.coef[q] == dst;
                if (seen) continue;
                const opf_field_s* src = leaf(s, l - 1, (int) k);
                CoefMap m{};
                for (int d = 0; d < 3; ++d) {
                    m.f0[d] = d < dim ? F.mesh->range.start[d] : 0;
                    m.c0[d] = d < dim ? C.mesh->range.start[d] : 0;
                    m.avg[d] = d < dim && dst->loc[d] == OPF_LOC_CENTER ? 1 : 0;
                }
                const Range r = common(dst->assignable, dst->local);
                if (r.count() > 0) {
                    coef_restrict_kernel<<<box_grid(r).grid, box_grid(r).block, 0, ctx().stream>>>(src->biased(src->cur), src->pitch1, src->pitch2, dst->biased(dst->cur),
                                                                                                 dst->pitch1, dst->pitch2, lr_of(r), m);
                    ctx().launches++;
                }
                if (int rc = field_update_padding(dst)) return rc;
            }
            if (C.c0) {// lhs_l(0) with homogeneous boundary data
                if (int rc = assign(C.x, "S<0>", {}, {0.0})) return rc;
                if (int rc = apply_lhs(s, C.x, C.c0, l, false, true)) return rc;
            }
        }
        OPF_CUDA(cudaGetLastError());
        return OPF_OK;
    }
}// namespace

extern "C" {

opf_solver_t opf_solver_create(opf_field_t target, const char* lhs_signature, const opf_field_t* lhs_fields, int n_lhs_fields,
                               const double* lhs_scalars, int n_lhs_scalars, unsigned unknown_mask, const opf_solver_params* params) {
    if (!target || !lhs_signature || !params) {
        fail(OPF_ERR_INVALID, "opf_solver_create: null argument");
        return nullptr;
    }
    if (n_lhs_fields > OPF_MAX_FIELDS - 1) {
        fail(OPF_ERR_UNSUPPORTED, "opf_solver_create: the equation has %d field leaves (max %d)", n_lhs_fields, OPF_MAX_FIELDS - 1);
        return nullptr;
    }
    if (require_device()) return nullptr;
    switch (params->type) {
        case OPF_SOLVER_PCG:
        case OPF_SOLVER_BICGSTAB:
        case OPF_SOLVER_GMRES:
        case OPF_SOLVER_FGMRES:
        case OPF_SOLVER_LGMRES:
        case OPF_SOLVER_JACOBI:
        case OPF_SOLVER_PFMG:
        case OPF_SOLVER_SMG: break;
        default: fail(OPF_ERR_UNSUPPORTED, "solver type %d is not implemented by the matrix-free engine", params->type); return nullptr;
    }
    auto* s = new opf_solver_s();
    s->target = target;
    s->lhs_sig = lhs_signature;
    for (int k = 0; k < n_lhs_fields; ++k) s->lhs_fields.push_back(((unknown_mask >> k) & 1u) ? nullptr : lhs_fields[k]);
    for (int k = 0; k < n_lhs_scalars; ++k) s->lhs_scalars.push_back(lhs_scalars[k]);
    s->mask = unknown_mask;
    s->params = *params;
    if (s->params.tol <= 0) s->params.tol = 1e-6;// HYPRE default tolerance
    if (s->params.max_iter <= 0) s->params.max_iter = 100;
    s->omega = 2.0 * target->dim / (2.0 * target->dim + 1.0);// weighted Jacobi: 2/3, 4/5, 6/7
    // fused residual kernel  r = b - lhs(x)  if that expression is compiled in: shift the leaf indices of lhs by one
    {
        std::string sh;
        const std::string& g = s->lhs_sig;
        for (size_t i = 0; i < g.size(); ++i) {
            if (g[i] == 'F' && i + 1 < g.size() && g[i + 1] == '<' && (i == 0 || !isalnum((unsigned char) g[i - 1]))) {
                size_t j = i + 2;
                int v = 0;
                while (j < g.size() && isdigit((unsigned char) g[j])) v = v * 10 + (g[j++] - '0');
                sh += "F<" + std::to_string(v + 1);
                i = j - 1;
            } else if (g[i] != ' ')
                sh.push_back(g[i]);
        }
        s->res_sig = "Sub<F<0>," + sh + ">";
        s->has_res_sig = opf_expr_is_registered(s->res_sig.c_str()) != 0;
        // fused weighted-Jacobi sweep  x <- x + w * dinv * (b - lhs(x)).  Leaves: F<0> = x (every unknown leaf of lhs maps to it, so
        // the kernel stages x once), F<1> = dinv, F<2> = b, coefficient fields of lhs follow from F<3>; S<0> = w, lhs scalars from S<1>
        std::string sm;
        std::vector<int> coef_leaf;// lhs leaf index of every coefficient field, in order of first appearance
        for (size_t i = 0; i < g.size(); ++i) {
            const bool leaf = (g[i] == 'F' || g[i] == 'S') && i + 1 < g.size() && g[i + 1] == '<' && (i == 0 || !isalnum((unsigned char) g[i - 1]));
            if (!leaf) {
                if (g[i] != ' ') sm.push_back(g[i]);
                continue;
            }
            size_t j = i + 2;
            int v = 0;
            while (j < g.size() && isdigit((unsigned char) g[j])) v = v * 10 + (g[j++] - '0');
            if (g[i] == 'S') sm += "S<" + std::to_string(v + 1);
            else if ((unknown_mask >> v) & 1u)
                sm += "F<0";
            else {
                size_t pos = 0;
                while (pos < coef_leaf.size() && coef_leaf[pos] != v) ++pos;
                if (pos == coef_leaf.size()) coef_leaf.push_back(v);
                sm += "F<" + std::to_string(3 + (int) pos);
            }
            i = j - 1;
        }
        s->smooth_coef_leaf = coef_leaf;
        s->smooth_sig = "Add<F<0>,Mul<S<0>,Mul<F<1>,Sub<F<2>," + sm + ">>>>";
        const int fused_on = opf_internal_opt(OPF_OPT_MG_FUSED);
        s->has_smooth_sig = fused_on && (int) coef_leaf.size() + 3 <= OPF_MAX_FIELDS && n_lhs_scalars + 1 <= OPF_MAX_SCALARS
                            && opf_expr_is_registered(s->smooth_sig.c_str()) != 0;
        // fused red-black half-sweep  x <- x + Par<c> * dinv * (b - lhs(x)): same leaf numbering, the colour mask in place of w
        for (int c = 0; c < 2; ++c) s->rb_sig[c] = "Add<F<0>,Mul<Par<" + std::to_string(c) + ">,Mul<F<1>,Sub<F<2>," + sm + ">>>>";
        s->has_rb_sig = fused_on && (int) coef_leaf.size() + 3 <= OPF_MAX_FIELDS && n_lhs_scalars + 1 <= OPF_MAX_SCALARS
                        && opf_expr_is_registered(s->rb_sig[0].c_str()) != 0 && opf_expr_is_registered(s->rb_sig[1].c_str()) != 0;
    }
    const Range w = common(target->assignable, target->local);
    s->pinned = params->pin_value != 0;
    s->pin_off = (long long) target->assignable.start[0] + (long long) target->assignable.start[1] * target->pitch1
                 + (long long) target->assignable.start[2] * target->pitch2;
    (void) w;
    if (build_levels(s)) {
        opf_solver_destroy(s);
        return nullptr;
    }
    s->X = clone_homogeneous(target, "kry.x");
    s->B = clone_homogeneous(target, "kry.b");
    s->R = clone_homogeneous(target, "kry.r");
    s->P = clone_homogeneous(target, "kry.p");
    s->Z = clone_homogeneous(target, "kry.z");
    s->Q = clone_homogeneous(target, "kry.q");
    s->E0 = opf_field_clone(target, "kry.e0");
    const bool bicg = params->type == OPF_SOLVER_BICGSTAB;
    if (bicg) {
        s->R0 = clone_homogeneous(target, "kry.r0");
        s->V = clone_homogeneous(target, "kry.v");
        s->S = clone_homogeneous(target, "kry.s");
        s->T = clone_homogeneous(target, "kry.t");
    }
    if (!s->X || !s->B || !s->R || !s->P || !s->Z || !s->Q || !s->E0) {
        opf_solver_destroy(s);
        return nullptr;
    }
    return s;
}

static void drop_graphs(opf_solver_s* s);
int opf_solver_levels(opf_solver_t s) { return s ? (int) s->lv.size() : -1; }


// CSRMatrixGenerator::generate (src/Core/Equation/CSRMatrixGenerator.hpp:55-138) for the matrix-free operator: rows and columns are
// the ranks of the target's assignable cells in x-fastest order (DS::MDRangeMapper), columns ascending inside a row, b = rhs - lhs(0 with
// the real boundary data) as the right-hand side, the LAST row replaced by the identity when pin_last is set (the CSR route pins the last
// row, :78,84-91).  The coefficients are read off the operator by coloured probing (the lattice of build_diag): one operator
// application per colour, each row picks the one column of that colour inside its footprint.  In STENCIL arithmetic the values are the
// reference's assembled ones bit for bit (tests/test_gpu_coefficients.py).  Single rank only.
int opf_solver_export_csr(opf_solver_t s, const char* rhs_signature, const opf_field_t* rhs_fields, int n_rhs_fields, const double* rhs_scalars,
                          int n_rhs_scalars, int pin_last, long long cap_nnz, int* ptr, int* col, double* val, double* rhs, long long* nnz_out) {
    if (!s || !ptr || !col || !val || !rhs || !rhs_signature) return fail(OPF_ERR_INVALID, "null argument");
    opf_field_s* t = s->target;
    if (t->n_ranks > 1) return fail(OPF_ERR_UNSUPPORTED, "opf_solver_export_csr: decomposed targets are not supported");
    const int dim = t->dim;
    const Range w = common(t->assignable, t->local);
    const long long nrows = w.count();
    if (nrows <= 0) return fail(OPF_ERR_INVALID, "empty assignable range");
    std::vector<std::array<int, 3>> taps;
    if (signature_unknown_taps(s->lhs_sig.c_str(), s->mask, taps) || taps.empty()) return fail(OPF_ERR_INVALID, "cannot derive the footprint of '%s'", s->lhs_sig.c_str());
    // colouring: any (a, b, c, M) separating the footprint differences; small problems simply use one colour per cell offset pattern
    std::vector<std::array<int, 3>> diff;
    for (const auto& p1 : taps)
        for (const auto& p2 : taps)
            if (p1 != p2) diff.push_back({p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]});
    long long per[3] = {0, 0, 0};
    int n[3] = {1, 1, 1};
    for (int d = 0; d < dim; ++d) {
        n[d] = w.end[d] - w.start[d];
        if (t->bc[d][0].type == OPF_BC_PERIODIC) per[d] = t->accessible.end[d] - t->accessible.start[d];
    }
    // periodic axes: only cells of the range itself are coloured (ghosts are images), so two cells also meet across the seam -- their
    // index difference is a footprint difference shifted by one period either way
    for (int d = 0; d < dim; ++d) {
        if (!per[d]) continue;
        const size_t nd = diff.size();
        for (size_t q = 0; q < nd; ++q)
            for (int sgn = -1; sgn <= 1; sgn += 2) {
                auto e = diff[q];
                e[d] += sgn * (int) per[d];
                diff.push_back(e);
            }
    }
    std::sort(diff.begin(), diff.end());
    diff.erase(std::unique(diff.begin(), diff.end()), diff.end());
    diff.erase(std::remove(diff.begin(), diff.end(), std::array<int, 3>{0, 0, 0}), diff.end());// a cell and itself
    int lat[4] = {0, 0, 0, 0};
    for (int M = (int) taps.size(); M <= 512 && !lat[3]; ++M)
        for (int a = 1; a < std::min(std::max(2, M), 17) && !lat[3]; ++a)
            for (int b = (dim >= 2 ? 1 : 0); b < (dim >= 2 ? M : 1) && !lat[3]; ++b)
                for (int c = (dim >= 3 ? 1 : 0); c < (dim >= 3 ? M : 1) && !lat[3]; ++c) {
                    bool ok = true;
                    for (const auto& d : diff) {
                        if ((((long long) a * d[0] + (long long) b * d[1] + (long long) c * d[2]) % M) == 0) {
                            ok = false;
                            break;
                        }
                    }
                    if (ok) lat[0] = a, lat[1] = b, lat[2] = c, lat[3] = M;
                }
    if (!lat[3]) return fail(OPF_ERR_UNSUPPORTED, "opf_solver_export_csr: no probe colouring found for this footprint / periodic extents");
    // b = rhs - lhs(0 with the real BC data), like opf_solver_solve
    s->pin_active = false;
    if (int rc = opf_assign_ex(s->B, OPF_OP_EQ, rhs_signature, rhs_fields, n_rhs_fields, rhs_scalars, n_rhs_scalars, OPF_ASSIGN_NO_PADDING)) return rc;
    if (int rc = assign(s->E0, "S<0>", {}, {0.0})) return rc;
    if (int rc = apply_lhs(s, s->E0, s->Q, 0, false, true)) return rc;
    if (int rc = assign(s->B, "Sub<F<0>,F<1>>", {s->B, s->Q}, {})) return rc;
    const opf_range wr = to_c(w);
    std::vector<double> bvals((size_t) nrows), q((size_t) nrows);
    if (int rc = opf_field_download(s->B, &wr, bvals.data())) return rc;
    // affine constant part (lhs terms without the unknown) is removed from every probe
    std::vector<double> c0((size_t) nrows, 0.0);
    {
        if (int rc = assign(s->Z, "S<0>", {}, {0.0})) return rc;
        if (int rc = apply_lhs(s, s->Z, s->Q, 0, false, true)) return rc;
        if (int rc = opf_field_download(s->Q, &wr, c0.data())) return rc;
    }
    auto rank_of = [&](const int g[3]) { return (long long) (g[0] - w.start[0]) + (long long) n[0] * ((g[1] - w.start[1]) + (long long) n[1] * (g[2] - w.start[2])); };
    auto colour = [&](const int g[3]) {
        long long v = ((long long) lat[0] * g[0] + (long long) lat[1] * g[1] + (long long) lat[2] * g[2]) % lat[3];
        return (int) (v < 0 ? v + lat[3] : v);
    };
    std::vector<std::vector<std::pair<int, double>>> rows((size_t) nrows);
    Mod3 mm{{1, 1, 1}, {lat[0], lat[1], lat[2], lat[3]}};
    const BoxGrid bg = box_grid(w);
    opf_field_s* X = s->lv[0].x;
    for (int cidx = 0; cidx < lat[3]; ++cidx) {
        color_fill_kernel<<<bg.grid, bg.block, 0, ctx().stream>>>(X->biased(X->cur), X->pitch1, X->pitch2, lr_of(w), dim, mm, cidx, 0, 0);
        ctx().launches++;
        if (int rc = apply_lhs(s, X, s->Q, 0, false, true)) return rc;
        if (int rc = opf_field_download(s->Q, &wr, q.data())) return rc;
        int g[3];
        for (g[2] = w.start[2]; g[2] < w.end[2]; ++g[2])
            for (g[1] = w.start[1]; g[1] < w.end[1]; ++g[1])
                for (g[0] = w.start[0]; g[0] < w.end[0]; ++g[0]) {
                    const long long r = rank_of(g);
                    const double v = q[(size_t) r] - c0[(size_t) r];
                    if (v == 0.0) continue;
                    // the column: the footprint cell of this colour (folded back into the range: periodic wrap, BC mirror)
                    int found = -1;
                    for (const auto& o : taps) {
                        int j[3] = {g[0] + o[0], g[1] + o[1], g[2] + o[2]};
                        bool inside = true;
                        for (int d = 0; d < dim; ++d) {
                            if (per[d]) {
                                while (j[d] < w.start[d]) j[d] += (int) per[d];
                                while (j[d] >= w.end[d]) j[d] -= (int) per[d];
                            } else if (j[d] < w.start[d] || j[d] >= w.end[d])
                                inside = false;
                        }
                        if (inside && colour(j) == cidx) {
                            found = (int) rank_of(j);
                            break;
                        }
                    }
                    if (found < 0) return fail(OPF_ERR_INVALID, "opf_solver_export_csr: probe response outside the footprint (row %lld, colour %d)", r, cidx);
                    rows[(size_t) r].push_back({found, v});
                }
    }
    if (int rc = assign(X, "S<0>", {}, {0.0})) return rc;
    long long nnz = 0;
    for (long long r = 0; r < nrows; ++r) {
        auto& row = rows[(size_t) r];
        if (pin_last && r == nrows - 1) {
            row.assign(1, {(int) r, 1.0});
            bvals[(size_t) r] = 0.0;
        }
        std::sort(row.begin(), row.end());
        ptr[r] = (int) nnz;
        for (const auto& e : row) {
            if (nnz >= cap_nnz) return fail(OPF_ERR_INVALID, "opf_solver_export_csr: more than %lld non-zeros", cap_nnz);
            col[nnz] = e.first;
            val[nnz] = e.second;
            ++nnz;
        }
        rhs[r] = bvals[(size_t) r];
    }
    ptr[nrows] = (int) nnz;
    if (nnz_out) *nnz_out = nnz;
    return OPF_OK;
}

int opf_solver_update(opf_solver_t s, const opf_field_t* lhs_fields, int n_lhs_fields, const double* lhs_scalars, int n_lhs_scalars) {
    if (!s) return fail(OPF_ERR_INVALID, "null solver");
    if (n_lhs_fields != (int) s->lhs_fields.size() || n_lhs_scalars != (int) s->lhs_scalars.size())
        return fail(OPF_ERR_INVALID, "opf_solver_update: leaf counts differ from opf_solver_create (%d/%d fields, %d/%d scalars)", n_lhs_fields,
                    (int) s->lhs_fields.size(), n_lhs_scalars, (int) s->lhs_scalars.size());
    bool changed = false;
    for (int k = 0; k < n_lhs_fields; ++k) {
        opf_field_s* f = ((s->mask >> k) & 1u) ? nullptr : lhs_fields[k];
        if (f != s->lhs_fields[k]) changed = true;
        s->lhs_fields[k] = f;
    }
    for (int k = 0; k < n_lhs_scalars; ++k) {
        if (lhs_scalars[k] != s->lhs_scalars[k]) changed = true;
        s->lhs_scalars[k] = lhs_scalars[k];
    }
    if (changed) {
        s->setup_done = false;// a static operator with new coefficients is set up again
        drop_graphs(s);       // scalars and field pointers are baked into the captured launches
    }
    return OPF_OK;
}

static void drop_graphs(opf_solver_s* s) {
    if (opfe::ctx().inited) cudaStreamSynchronize(opfe::ctx().stream);
    for (auto& g : s->vgraphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    s->vgraphs.clear();
    for (auto& g : s->loops)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    s->loops.clear();
}

int opf_solver_destroy(opf_solver_t s) {
    if (!s) return OPF_OK;
    drop_graphs(s);
    if (s->ks) cudaFree(s->ks);
    if (s->ctl_dev) cudaFree(s->ctl_dev);
    if (s->ctl_host) cudaFreeHost(s->ctl_host);
    for (auto& L : s->lv) free_level_fields(L);
    for (opf_field_s* f : {s->X, s->B, s->R, s->P, s->Z, s->Q, s->R0, s->V, s->S, s->T, s->E0, s->C0})
        if (f) opf_field_destroy(f);
    for (opf_field_s* f : s->gm_v) opf_field_destroy(f);
    delete s;
    return OPF_OK;
}


// ---- device-resident PCG (see the kernels above).  One iteration, launched into the engine stream (plainly, or into a capture):
//   q = A p ; pq = p.q ; x += (rz/pq) p ; r -= (rz/pq) q ; rr = r.r ; [control] ; z = M^-1 r ; rzn = r.z ; p = z + (rzn/rz) p ; rz = rzn
// geometry of the streaming Krylov kernels over box w of fields cloned from the target (same pitches, same alignment)
static int krylov_blocks(const Range& w) {
    const long long nseg = (w.end[0] - w.start[0] + KSEG - 1) / KSEG;
    const long long items = nseg * (w.end[1] - w.start[1]) * (long long) (w.end[2] - w.start[2]);
    return (int) std::max<long long>(1, std::min<long long>(items, 8LL * ctx().sm_count));
}
static int krylov_vec(std::initializer_list<const opf_field_s*> fs, const Range& w) {
    for (const opf_field_s* f : fs) {
        const double* p0 = f->biased(f->cur) + ((long long) w.start[0] + (long long) w.start[1] * f->pitch1 + (long long) w.start[2] * f->pitch2);
        if ((reinterpret_cast<uintptr_t>(p0) & 15) || (f->pitch1 & 1) || (f->pitch2 & 1)) return 0;
    }
    return 1;
}
// slot = sum over w of a.b, device-resident (two launches, no host synchronisation)
static int krylov_dot(opf_field_s* a, opf_field_s* b, const Range& w, double* slot) {
    Context& c = ctx();
    if (a->pitch1 != b->pitch1 || a->pitch2 != b->pitch2) return dot_device(a, b, w, slot);
    const int nb = krylov_blocks(w);
    if (nb > c.red_cap - 8) return dot_device(a, b, w, slot);
    dot_fast_kernel<<<nb, 256, 0, c.stream>>>(a->biased(a->cur), b->biased(b->cur), a->pitch1, a->pitch2, lr_of(w), c.red_buf, krylov_vec({a, b}, w));
    fold_kernel<<<1, 256, 0, c.stream>>>(c.red_buf, nb, slot);
    c.launches += 2;
    return OPF_OK;
}

static int pcg_body(opf_solver_s* s, const Range& w, cudaGraphConditionalHandle handle, int use_cond) {
    Context& c = ctx();
    const bool dist = s->lv[0].dist && comm_active();
    if (int rc = apply_lhs(s, s->P, s->Q, 0)) return rc;
    if (int rc = krylov_dot(s->P, s->Q, w, s->ks + KS_PQ)) return rc;
    if (dist)
        if (int rc = comm_allreduce_device(s->ks + KS_PQ, 1, OPF_RED_SUM, c.stream)) return rc;
    const int nb = std::min(krylov_blocks(w), c.red_cap - 8);
    opf_field_s *X = s->X, *P = s->P, *R = s->R, *Q = s->Q;
    pcg_update_kernel<<<nb, 256, 0, c.stream>>>(X->biased(X->cur), P->biased(P->cur), R->biased(R->cur), Q->biased(Q->cur), X->pitch1, X->pitch2, lr_of(w), s->ks,
                                                c.red_buf, krylov_vec({X, P, R, Q}, w));
    if (!dist) {
        pcg_control_kernel<<<1, 256, 0, c.stream>>>(c.red_buf, nb, s->ks, s->ctl_dev, handle, use_cond, 0);
        c.launches += 2;
    } else {
        pcg_control_kernel<<<1, 256, 0, c.stream>>>(c.red_buf, nb, s->ks, s->ctl_dev, handle, 0, 1);
        if (int rc = comm_allreduce_device(s->ks + KS_RR, 1, OPF_RED_SUM, c.stream)) return rc;
        pcg_control_kernel<<<1, 256, 0, c.stream>>>(nullptr, 0, s->ks, s->ctl_dev, handle, 0, 0);
        c.launches += 3;
    }
    if (int rc = precondition_pinned(s, s->R, s->Z)) return rc;
    if (int rc = krylov_dot(s->R, s->Z, w, s->ks + KS_RZN)) return rc;
    if (dist)
        if (int rc = comm_allreduce_device(s->ks + KS_RZN, 1, OPF_RED_SUM, c.stream)) return rc;
    opf_field_s* Z = s->Z;
    pcg_direction_kernel<<<krylov_blocks(w), 256, 0, c.stream>>>(P->biased(P->cur), Z->biased(Z->cur), P->pitch1, P->pitch2, lr_of(w), s->ks, krylov_vec({P, Z}, w));
    pcg_shift_kernel<<<1, 1, 0, c.stream>>>(s->ks);
    c.launches += 2;
    OPF_CUDA(cudaGetLastError());
    return OPF_OK;
}

static unsigned loop_parity(opf_solver_s* s) {
    unsigned pp = (unsigned) s->Z->cur | coef_parity(s);
    for (size_t lv = 1; lv < s->lv.size(); ++lv) pp |= (unsigned) s->lv[lv].x->cur << lv;
    return pp;
}
static void loop_apply_parity(opf_solver_s* s, unsigned pp) {
    s->Z->cur = (int) (pp & 1u);
    for (size_t lv = 1; lv < s->lv.size(); ++lv) s->lv[lv].x->cur = (int) ((pp >> lv) & 1u);
}

// PCG on the current operator with device-resident scalars.  rel0 / rr0: relative residual and r.r of the starting iterate (already
// computed by the caller, s->R holds that residual).
static int run_pcg_device(opf_solver_s* s, const Range& w, double bnorm, double tol, int maxit, double rr0, int* iters_io, double* rel_io) {
    Context& c = ctx();
    const bool dist = s->lv[0].dist && comm_active();
    if (!s->ks) {
        OPF_CUDA(cudaMalloc(&s->ks, sizeof(double) * KS_COUNT));
        OPF_CUDA(cudaMalloc(&s->ctl_dev, sizeof(PcgCtl)));
        OPF_CUDA(cudaMallocHost(&s->ctl_host, sizeof(PcgCtl)));
    }
    double init[KS_COUNT] = {0};
    init[KS_RR] = rr0, init[KS_BNORM] = bnorm, init[KS_TOL] = tol, init[KS_PQ] = 1.0;
    OPF_CUDA(cudaMemcpyAsync(s->ks, init, sizeof init, cudaMemcpyHostToDevice, c.stream));
    *s->ctl_host = PcgCtl{*iters_io, maxit, 1, 0, *rel_io};
    OPF_CUDA(cudaMemcpyAsync(s->ctl_dev, s->ctl_host, sizeof(PcgCtl), cudaMemcpyHostToDevice, c.stream));
    // z = M^-1 r ; p = z ; rz = r.z
    if (int rc = precondition_pinned(s, s->R, s->Z)) return rc;
    if (int rc = assign(s->P, "F<0>", {s->Z}, {})) return rc;
    if (int rc = krylov_dot(s->R, s->Z, w, s->ks + KS_RZ)) return rc;
    if (dist)
        if (int rc = comm_allreduce_device(s->ks + KS_RZ, 1, OPF_RED_SUM, c.stream)) return rc;
    auto fetch = [&]() -> int {
        OPF_CUDA(cudaMemcpyAsync(s->ctl_host, s->ctl_dev, sizeof(PcgCtl), cudaMemcpyDeviceToHost, c.stream));
        OPF_CUDA(cudaStreamSynchronize(c.stream));
        return OPF_OK;
    };
    const bool want_graph = opf_internal_opt(OPF_OPT_GRAPHS) && opf_internal_opt(OPF_OPT_FUSED_KRYLOV) >= 2 && !dist;
    // the first iteration always runs as plain launches: every lazy allocation (twin buffers, reduction scratch, kernel attributes)
    // happens here, outside any capture
    if (int rc = pcg_body(s, w, 0, 0)) return rc;
    if (int rc = fetch()) return rc;
    while (s->ctl_host->cont) {
        if (want_graph) {
            const unsigned parity = loop_parity(s);
            opf_solver_s::LoopGraph* g = nullptr;
            for (auto& e : s->loops)
                if (e.parity == parity && e.mode == c.mode && e.precond == s->params.precond && e.pin_active == (int) s->pin_active) g = &e;
            if (!g) {
                // capture the iteration as the body of a WHILE node: the control kernel sets the loop condition on the device
                cudaGraph_t graph = nullptr;
                OPF_CUDA(cudaGraphCreate(&graph, 0));
                cudaGraphConditionalHandle handle;
                OPF_CUDA(cudaGraphConditionalHandleCreate(&handle, graph, 1, cudaGraphCondAssignDefault));
                cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
                np.conditional.handle = handle;
                np.conditional.type = cudaGraphCondTypeWhile;
                np.conditional.size = 1;
                cudaGraphNode_t node;
                cudaError_t ce = cudaGraphAddNode(&node, graph, nullptr, 0, &np);
                if (ce == cudaSuccess) ce = cudaStreamBeginCaptureToGraph(c.stream, np.conditional.phGraph_out[0], nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
                if (ce != cudaSuccess) {
                    cudaGraphDestroy(graph);
                    cudaGetLastError();
                    s->loops.push_back({parity, parity, c.mode, s->params.precond, (int) s->pin_active, nullptr, 0});// marks "do not try again"
                    g = &s->loops.back();
                } else {
                    const long long l0 = c.launches;
                    s->in_loop_capture = true;
                    const int rc = pcg_body(s, w, handle, 1);
                    s->in_loop_capture = false;
                    cudaGraph_t done = nullptr;
                    ce = cudaStreamEndCapture(c.stream, &done);
                    const unsigned after = loop_parity(s);
                    loop_apply_parity(s, parity);
                    const long long nl = c.launches - l0;
                    c.launches = l0;
                    cudaGraphExec_t exec = nullptr;
                    if (!rc && ce == cudaSuccess) ce = cudaGraphInstantiate(&exec, graph, 0);
                    cudaGraphDestroy(graph);
                    if (rc) return rc;
                    if (ce != cudaSuccess) {
                        cudaGetLastError();
                        exec = nullptr;
                    }
                    s->loops.push_back({parity, after, c.mode, s->params.precond, (int) s->pin_active, exec, nl});
                    g = &s->loops.back();
                }
            }
            if (g->exec) {
                const int before = s->ctl_host->iters;
                OPF_CUDA(cudaGraphLaunch(g->exec, c.stream));
                if (int rc = fetch()) return rc;
                const int ran = s->ctl_host->iters - before + (s->ctl_host->breakdown ? 1 : 0);
                if (ran > 0) loop_apply_parity(s, g->parity_after);
                c.launches += g->launches * std::max(ran, 1);
                break;// the loop ran to its end on the device
            }
        }
        if (int rc = pcg_body(s, w, 0, 0)) return rc;
        if (int rc = fetch()) return rc;
    }
    *iters_io = s->ctl_host->iters;
    *rel_io = s->ctl_host->rel;
    return OPF_OK;
}

// one Krylov / stationary run on the current operator (s->pin_active), right-hand side s->B and iterate s->X
static int run_iteration(opf_solver_s* s, int type, const Range& w, double bnorm, double tol, int maxit, int* iters_io, double* rel_out) {
    auto& L0 = s->lv[0];
    double rnorm2 = 0;
    int iters = *iters_io;
    if (int rc = residual(s, s->X, s->B, s->R, s->Q, 0)) return rc;
    if (int rc = dot(s, s->R, s->R, w, &rnorm2)) return rc;
    double rel = std::sqrt(rnorm2) / bnorm;
    if (type == OPF_SOLVER_PCG && opf_internal_opt(OPF_OPT_FUSED_KRYLOV) && rel > tol && iters < maxit) {
        if (int rc = run_pcg_device(s, w, bnorm, tol, maxit, rnorm2, &iters, &rel)) return rc;
    } else if (type == OPF_SOLVER_PCG) {
        // preconditioned conjugate gradients, convergence on ||r||2 / ||b||2 (host-driven reference implementation of the loop)
        double rz = 0, rz_new = 0, pq = 0;
        if (rel > tol) {
            if (int rc = precondition_pinned(s, s->R, s->Z)) return rc;
            if (int rc = assign(s->P, "F<0>", {s->Z}, {})) return rc;
            if (int rc = dot(s, s->R, s->Z, w, &rz)) return rc;
        }
        while (rel > tol && iters < maxit) {
            if (int rc = apply_lhs(s, s->P, s->Q, 0)) return rc;
            if (int rc = dot(s, s->P, s->Q, w, &pq)) return rc;
            if (pq == 0.0) break;
            const double alpha = rz / pq;
            if (int rc = assign(s->X, "Add<F<0>,Mul<S<0>,F<1>>>", {s->X, s->P}, {alpha})) return rc;
            if (int rc = assign(s->R, "Add<F<0>,Mul<S<0>,F<1>>>", {s->R, s->Q}, {-alpha})) return rc;
            if (int rc = dot(s, s->R, s->R, w, &rnorm2)) return rc;
            ++iters;
            rel = std::sqrt(rnorm2) / bnorm;
            if (rel <= tol) break;
            if (int rc = precondition_pinned(s, s->R, s->Z)) return rc;
            if (int rc = dot(s, s->R, s->Z, w, &rz_new)) return rc;
            const double beta = rz_new / rz;
            rz = rz_new;
            if (int rc = assign(s->P, "Add<F<0>,Mul<S<0>,F<1>>>", {s->Z, s->P}, {beta})) return rc;
        }
    } else if (type == OPF_SOLVER_JACOBI || type == OPF_SOLVER_PFMG || type == OPF_SOLVER_SMG) {
        // stationary iteration: weighted Jacobi sweeps or V-cycles on the residual equation
        while (rel > tol && iters < maxit) {
            if (type == OPF_SOLVER_JACOBI) {
                if (int rc = assign(s->X, "Add<F<0>,Mul<S<0>,Mul<F<1>,F<2>>>>", {s->X, L0.dinv, s->R}, {s->omega})) return rc;
            } else {
                if (int rc = assign(L0.b, "F<0>", {s->R}, {})) return rc;
                if (int rc = vcycle(s, 0, true)) return rc;
                if (int rc = assign(s->X, "Add<F<0>,F<1>>", {s->X, L0.x}, {})) return rc;
            }
            if (s->pin_active) poke(s, s->X, s->pin_off, 0.0);
            if (int rc = residual(s, s->X, s->B, s->R, s->Q, 0)) return rc;
            if (int rc = dot(s, s->R, s->R, w, &rnorm2)) return rc;
            ++iters;
            rel = std::sqrt(rnorm2) / bnorm;
        }
    } else if (type == OPF_SOLVER_GMRES || type == OPF_SOLVER_FGMRES || type == OPF_SOLVER_LGMRES) {
        // restarted GMRES(k), right-preconditioned (HYPRE_StructGMRES*, StructSolverGMRES.hpp:20-96; hypre's default k_dim = 5), modified
        // Gram-Schmidt, Givens rotations on the host.  The preconditioner is a fixed linear operator here (Jacobi or a V-cycle), so the
        // flexible variant (FGMRES) coincides with it; LGMRES requests run as GMRES with the augmentation vectors added to k.
        const int k = std::max(1, std::min(s->params.k_dim > 0 ? s->params.k_dim : 5, 50));
        if ((int) s->gm_v.size() < k + 1) {
            for (int i = (int) s->gm_v.size(); i < k + 1; ++i) {
                opf_field_s* v = clone_homogeneous(s->target, "kry.gmres.v");
                if (!v) return OPF_ERR_CUDA;
                s->gm_v.push_back(v);
            }
        }
        std::vector<double> H((size_t) (k + 1) * k, 0.0), cs(k, 0.0), sn(k, 0.0), g(k + 1, 0.0), y(k, 0.0);
        auto h = [&](int i, int j) -> double& { return H[(size_t) i * k + j]; };
        while (rel > tol && iters < maxit) {
            const double beta = std::sqrt(rnorm2);
            if (beta == 0.0) break;
            if (int rc = assign(s->gm_v[0], "Mul<S<0>,F<0>>", {s->R}, {1.0 / beta})) return rc;
            std::fill(g.begin(), g.end(), 0.0);
            g[0] = beta;
            int j = 0;
            bool breakdown = false;
            for (; j < k && iters < maxit; ++j) {
                if (int rc = precondition_pinned(s, s->gm_v[j], s->Z)) return rc;// z = M^-1 v_j
                if (int rc = apply_lhs(s, s->Z, s->Q, 0)) return rc;             // w = A z
                for (int i = 0; i <= j; ++i) {
                    double hij = 0;
                    if (int rc = dot(s, s->Q, s->gm_v[i], w, &hij)) return rc;
                    h(i, j) = hij;
                    if (int rc = assign(s->Q, "Add<F<0>,Mul<S<0>,F<1>>>", {s->Q, s->gm_v[i]}, {-hij})) return rc;
                }
                double wn2 = 0;
                if (int rc = dot(s, s->Q, s->Q, w, &wn2)) return rc;
                const double wn = std::sqrt(wn2);
                h(j + 1, j) = wn;
                if (wn != 0.0) {
                    if (int rc = assign(s->gm_v[j + 1], "Mul<S<0>,F<0>>", {s->Q}, {1.0 / wn})) return rc;
                } else
                    breakdown = true;// lucky breakdown: the Krylov space contains the solution
                for (int i = 0; i < j; ++i) {// previous rotations on the new column
                    const double t = cs[i] * h(i, j) + sn[i] * h(i + 1, j);
                    h(i + 1, j) = -sn[i] * h(i, j) + cs[i] * h(i + 1, j);
                    h(i, j) = t;
                }
                const double den = std::hypot(h(j, j), h(j + 1, j));
                cs[j] = den != 0.0 ? h(j, j) / den : 1.0;
                sn[j] = den != 0.0 ? h(j + 1, j) / den : 0.0;
                h(j, j) = cs[j] * h(j, j) + sn[j] * h(j + 1, j);
                h(j + 1, j) = 0.0;
                g[j + 1] = -sn[j] * g[j];
                g[j] = cs[j] * g[j];
                ++iters;
                rel = std::fabs(g[j + 1]) / bnorm;
                if (rel <= tol || breakdown) {
                    ++j;
                    break;
                }
            }
            // y = H^-1 g ; x += M^-1 (V y)
            for (int i = j - 1; i >= 0; --i) {
                double acc = g[i];
                for (int c = i + 1; c < j; ++c) acc -= h(i, c) * y[c];
                y[i] = h(i, i) != 0.0 ? acc / h(i, i) : 0.0;
            }
            if (j > 0) {
                if (int rc = assign(s->Q, "Mul<S<0>,F<0>>", {s->gm_v[0]}, {y[0]})) return rc;
                for (int i = 1; i < j; ++i)
                    if (int rc = assign(s->Q, "Add<F<0>,Mul<S<0>,F<1>>>", {s->Q, s->gm_v[i]}, {y[i]})) return rc;
                if (int rc = precondition_pinned(s, s->Q, s->Z)) return rc;
                if (int rc = assign(s->X, "Add<F<0>,F<1>>", {s->X, s->Z}, {})) return rc;
            }
            // true residual for the restart.  On convergence the Givens estimate IS the residual norm of the right-preconditioned
            // iteration (exact arithmetic) and is what gets reported; OPF_GMRES_VERIFY=1 recomputes b - A x there as well (one more
            // operator application per solve, which for the 2-3 iteration momentum solves is a quarter of the work).
            static const int verify = getenv("OPF_GMRES_VERIFY") ? atoi(getenv("OPF_GMRES_VERIFY")) : 0;
            if (rel <= tol && !verify && rel > 1e-13) break;
            if (int rc = residual(s, s->X, s->B, s->R, s->Q, 0)) return rc;
            if (int rc = dot(s, s->R, s->R, w, &rnorm2)) return rc;
            rel = std::sqrt(rnorm2) / bnorm;
            if (breakdown || j == 0) break;
        }
    } else {
        // right-preconditioned BiCGSTAB (StructSolverBiCGSTAB.hpp)
        double rho = 1, alpha = 1, omega = 1, rho_new = 0, r0v = 0, ts = 0, tt = 0;
        if (int rc = assign(s->R0, "F<0>", {s->R}, {})) return rc;
        if (int rc = assign(s->P, "S<0>", {}, {0.0})) return rc;
        if (int rc = assign(s->V, "S<0>", {}, {0.0})) return rc;
        while (rel > tol && iters < maxit) {
            if (int rc = dot(s, s->R0, s->R, w, &rho_new)) return rc;
            if (rho_new == 0.0) break;
            const double beta = (rho_new / rho) * (alpha / omega);
            rho = rho_new;
            // p = r + beta (p - omega v)
            if (int rc = assign(s->P, "Add<F<0>,Mul<S<0>,F<1>>>", {s->P, s->V}, {-omega})) return rc;
            if (int rc = assign(s->P, "Add<F<0>,Mul<S<0>,F<1>>>", {s->R, s->P}, {beta})) return rc;
            if (int rc = precondition_pinned(s, s->P, s->Z)) return rc;// z = M^-1 p
            if (int rc = apply_lhs(s, s->Z, s->V, 0)) return rc;
            if (int rc = dot(s, s->R0, s->V, w, &r0v)) return rc;
            if (r0v == 0.0) break;
            alpha = rho / r0v;
            if (int rc = assign(s->S, "Add<F<0>,Mul<S<0>,F<1>>>", {s->R, s->V}, {-alpha})) return rc;
            if (int rc = assign(s->X, "Add<F<0>,Mul<S<0>,F<1>>>", {s->X, s->Z}, {alpha})) return rc;
            if (int rc = dot(s, s->S, s->S, w, &rnorm2)) return rc;
            if (std::sqrt(rnorm2) / bnorm <= tol) {
                ++iters;
                rel = std::sqrt(rnorm2) / bnorm;
                break;
            }
            if (int rc = precondition_pinned(s, s->S, s->Z)) return rc;// z = M^-1 s
            if (int rc = apply_lhs(s, s->Z, s->T, 0)) return rc;
            if (int rc = dot(s, s->T, s->S, w, &ts)) return rc;
            if (int rc = dot(s, s->T, s->T, w, &tt)) return rc;
            if (tt == 0.0) break;
            omega = ts / tt;
            if (int rc = assign(s->X, "Add<F<0>,Mul<S<0>,F<1>>>", {s->X, s->Z}, {omega})) return rc;
            if (int rc = assign(s->R, "Add<F<0>,Mul<S<0>,F<1>>>", {s->S, s->T}, {-omega})) return rc;
            if (int rc = dot(s, s->R, s->R, w, &rnorm2)) return rc;
            ++iters;
            rel = std::sqrt(rnorm2) / bnorm;
            if (omega == 0.0) break;
        }
    }
    *iters_io = iters;
    *rel_out = rel;
    return OPF_OK;
}

int opf_solver_solve(opf_solver_t s, const char* rhs_signature, const opf_field_t* rhs_fields, int n_rhs_fields, const double* rhs_scalars,
                     int n_rhs_scalars, opf_solve_state* state) {
    if (!s || !rhs_signature) return fail(OPF_ERR_INVALID, "null argument");
    opf_field_s* t = s->target;
    const Range w = common(t->assignable, t->local);
    // ---- setup (once when static_mat, else every solve -- the reference re-creates the HYPRE solver every time)
    // constant part of an affine lhs (multigrid levels only exist for pure operators, which have none).  Measured on EVERY solve,
    // static_mat or not: a coefficient field's contents may change between solves without its handle changing, and the reference
    // rebuilds the bias every time as well (generateb, HYPREEqnSolveHandler.hpp:199) -- one operator application.
    bool c0_measured = false, c0_not_needed = false;
    {
        s->pin_active = false;
        s->affine = false;
        bool pure = true;
        for (size_t k = 0; k < s->lhs_fields.size(); ++k)
            if (!((s->mask >> k) & 1u)) pure = false;
        c0_not_needed = pure && s->lhs_scalars.empty();
        if (!pure || !s->lhs_scalars.empty()) {
            c0_measured = true;
            if (!s->C0 && !(s->C0 = clone_homogeneous(t, "kry.c0"))) return OPF_ERR_CUDA;
            if (int rc = assign(s->Z, "S<0>", {}, {0.0})) return rc;
            if (int rc = apply_lhs(s, s->Z, s->C0, 0, false)) return rc;
            double cmax = 0;
            opf_field_t F[1] = {s->C0};
            opf_range cr = to_c(w);
            if (int rc = opf_reduce(OPF_RED_ABSMAX, "F<0>", F, 1, nullptr, 0, &cr, &cmax)) return rc;
            if (s->lv[0].dist)
                if (int rc = opf_comm_allreduce(&cmax, 1, OPF_RED_MAX)) return rc;
            s->affine = cmax != 0.0;
        }
    }
    // A non-static operator (the reference re-assembles and re-sets-up HYPRE on every solve) only needs its PRECONDITIONER data here --
    // the diagonal behind Jacobi / the smoothers; the operator itself is always applied with the current coefficient fields.  That
    // diagonal is refreshed every OPF_DIAG_REFRESH-th solve (default 8): a lagged diagonal changes iteration counts marginally, never
    // the converged solution, and its extraction costs one operator application per probe colour.
    static const int diag_refresh = getenv("OPF_DIAG_REFRESH") ? std::max(1, atoi(getenv("OPF_DIAG_REFRESH"))) : 8;
    const bool refresh = !s->params.static_mat && (s->solves % diag_refresh == 0);
    s->solves++;
    if (!s->setup_done || refresh) {
        s->pin_active = false;
        // a multigrid request on an operator that cannot be coarsened (coefficient fields, decomposed target) degrades to its
        // level-0 smoother, i.e. Jacobi: the diagonal is needed then as well
        const bool wants_mg = s->params.precond == OPF_SOLVER_PFMG || s->params.precond == OPF_SOLVER_SMG || s->params.type == OPF_SOLVER_PFMG
                              || s->params.type == OPF_SOLVER_SMG;
        const bool need_diag = s->mg || wants_mg || s->params.precond == OPF_SOLVER_JACOBI || s->params.type == OPF_SOLVER_JACOBI || s->pinned;
        if (need_diag)
            if (int rc = build_diag(s, 0)) return rc;
        if (s->coef_mg) {
            // Is the hierarchy worth its cost?  For  a I + (convection) - b Laplacian  the row sum over the diagonal, r = |A.1| / |d|, is
            // a / (a + 2 dim b / h^2): Jacobi-preconditioned Krylov sees a condition number of about (2 - r) / r.  r >= 1/4 everywhere
            // (kappa <= 7, a handful of iterations) -> the V-cycle stops at its level-0 smoother, like round 1's behaviour, and the
            // coarse levels are left alone.
            static const int coef_opt = getenv("OPF_MG_COEF") ? atoi(getenv("OPF_MG_COEF")) : 1;
            auto& L0 = s->lv[0];
            double rmin = 0;
            if (int rc = assign(L0.x, "S<0>", {}, {1.0})) return rc;
            if (int rc = apply_lhs(s, L0.x, L0.q, 0, false)) return rc;
            if (int rc = assign(L0.q, "Mul<F<0>,F<1>>", {L0.q, L0.dinv}, {})) return rc;
            opf_field_t F[1] = {L0.q};
            opf_range cr = to_c(w);
            if (int rc = opf_reduce(OPF_RED_MIN, "Abs<F<0>>", F, 1, nullptr, 0, &cr, &rmin)) return rc;
            if (int rc = assign(L0.x, "S<0>", {}, {0.0})) return rc;
            const bool bypass = coef_opt < 2 && rmin >= 0.25;
            if (bypass != s->mg_bypass) drop_graphs(s);// the captured V-cycle / loop bodies contain the other shape
            s->mg_bypass = bypass;
            if (!bypass)
                if (int rc = restrict_coefs(s)) return rc;
        }
        if (need_diag && s->mg && !s->mg_bypass)
            for (int l = 1; l < (int) s->lv.size(); ++l)
                if (int rc = build_diag(s, l)) return rc;
        if (s->pinned) {
            // is the un-pinned operator singular with the constants as null space?  ||A.1||_inf vs ||diag||_inf
            auto& L0 = s->lv[0];
            double a1 = 0, dmin = 0;
            if (int rc = assign(L0.x, "S<0>", {}, {1.0})) return rc;
            if (int rc = apply_lhs(s, L0.x, L0.q, 0, false)) return rc;
            opf_field_t F[1] = {L0.q};
            opf_range cr = to_c(w);
            if (int rc = opf_reduce(OPF_RED_ABSMAX, "F<0>", F, 1, nullptr, 0, &cr, &a1)) return rc;
            F[0] = L0.dinv;
            if (int rc = opf_reduce(OPF_RED_ABSMAX, "F<0>", F, 1, nullptr, 0, &cr, &dmin)) return rc;// max |1/d| = 1 / min |d|
            if (s->lv[0].dist) {
                double v[2] = {a1, dmin};
                if (int rc = opf_comm_allreduce(v, 2, OPF_RED_MAX)) return rc;
                a1 = v[0], dmin = v[1];
            }
            if (int rc = assign(L0.x, "S<0>", {}, {0.0})) return rc;
            s->singular = dmin > 0 && a1 * dmin <= 1e-8;
        }
        s->setup_done = true;
    }
    // ---- b = rhs - lhs(e = 0 with the real boundary data)          (generateAb / generateb :125-179: b = -bias)
    s->pin_active = false;
    if (int rc = opf_assign_ex(s->B, OPF_OP_EQ, rhs_signature, rhs_fields, n_rhs_fields, rhs_scalars, n_rhs_scalars, OPF_ASSIGN_NO_PADDING)) return rc;
    // A target whose boundary data is all zero (periodic, homogeneous Dirichlet / Neumann: every increment field du, dv, dp of the
    // projection schemes) has lhs(0 with the real BCs) == lhs(0 with homogeneous BCs) == C0, measured above -- or exactly 0 for an
    // operator without e-free terms: the operator application is skipped.
    bool zero_bc = true;
    for (int d = 0; d < t->dim; ++d)
        for (int sd = 0; sd < 2; ++sd) {
            const auto& bc = t->bc[d][sd];
            if (bc.type != OPF_BC_PERIODIC && (bc.value != 0.0 || bc.face_dev)) zero_bc = false;
        }
    static const int skip_e0 = getenv("OPF_SOLVER_SKIP_E0") ? atoi(getenv("OPF_SOLVER_SKIP_E0")) : 1;
    if (zero_bc && skip_e0 && c0_measured) {
        if (s->affine)
            if (int rc = assign(s->B, "Sub<F<0>,F<1>>", {s->B, s->C0}, {})) return rc;
    } else if (zero_bc && skip_e0 && c0_not_needed) {
        // pure operator, no scalars, zero boundary data: lhs(0) = 0
    } else {
        if (int rc = assign(s->E0, "S<0>", {}, {0.0})) return rc;
        if (int rc = apply_lhs(s, s->E0, s->Q, 0, false, true)) return rc;// E0 keeps the target's real BCs: its ghosts carry the boundary data
        if (int rc = assign(s->B, "Sub<F<0>,F<1>>", {s->B, s->Q}, {})) return rc;
    }
    // ---- x0 = current target values (initx :119-123)
    if (int rc = assign(s->X, "F<0>", {t}, {})) return rc;
    double bnorm2 = 0;
    if (int rc = dot(s, s->B, s->B, w, &bnorm2)) return rc;
    const double bnorm = std::sqrt(bnorm2);
    int iters = 0;
    double rel = 0;
    auto finish = [&](int rc_in) {
        if (rc_in) return rc_in;
        static const bool dbg = getenv("OPF_SOLVER_DEBUG") != nullptr;
        if (dbg) fprintf(stderr, "[opf_solver] target=%s lhs=%s type=%d precond=%d mg=%d coef_mg=%d bypass=%d levels=%d affine=%d pinned=%d singular=%d |b|=%.6e iters=%d rel=%.3e\n", t->name.c_str(), s->lhs_sig.substr(0, 40).c_str(), s->params.type, s->params.precond, (int) s->mg, (int) s->coef_mg, (int) s->mg_bypass, (int) s->lv.size(), (int) s->affine, (int) s->pinned, (int) s->singular, bnorm, iters, rel);
        if (state) {
            state->niter = iters;
            state->relerr = rel;
            state->abserr = rel * bnorm;
        }
        // returnValues (:181-188): x -> target, target.updatePadding()
        return opf_field_assign_field(t, OPF_OP_EQ, s->X);
    };
    if (bnorm == 0.0) {
        if (int rc = assign(s->X, "S<0>", {}, {0.0})) return rc;
        return finish(OPF_OK);
    }
    const double tol = s->params.tol;
    const int maxit = s->params.max_iter;
    const int type = s->params.type;
    if (s->pinned && s->singular) {
        // Phase 1: the consistent singular system on the mean-free subspace (un-pinned operator, b made mean-free), whose
        // solution shifted to x[pin] = 0 IS the reference's pinned solution whenever b is in the range of the operator.
        if (int rc = project_mean(s, s->B, w, s->lv[0].g, s->lv[0].dist)) return rc;
        if (int rc = run_iteration(s, type, w, bnorm, tol, maxit, &iters, &rel)) return rc;
        {// x -= x[pin]: the pinned cell's value, known to its owner, reaches every rank through a one-element sum
            double* scal = ctx().red_buf + ctx().red_cap - 1;
            if (s->lv[0].owns_pin) copy_cell_kernel<<<1, 1, 0, ctx().stream>>>(scal - s->pin_off, s->X->biased(s->X->cur), s->pin_off);
            else
                OPF_CUDA(cudaMemsetAsync(scal, 0, sizeof(double), ctx().stream));
            if (s->lv[0].dist)
                if (int rc = comm_allreduce_device(scal, 1, OPF_RED_SUM, ctx().stream)) return rc;
            sub_mean_kernel<<<box_grid(w).grid, box_grid(w).block, 0, ctx().stream>>>(s->X->biased(s->X->cur), s->X->pitch1, s->X->pitch2, lr_of(w), scal, 1.0);
            ctx().launches += 2;
        }
        // Phase 2: the reference's pinned system proper (row of the first assignable cell = identity, rhs 0).  For a consistent
        // b it is already converged; an inconsistent b is polished with Jacobi-preconditioned iterations on the pinned operator.
        s->pin_active = true;
        poke(s, s->B, s->pin_off, 0.0);
        const int saved_pre = s->params.precond;
        s->params.precond = OPF_SOLVER_JACOBI;
        const int ptype = (type == OPF_SOLVER_PCG || type == OPF_SOLVER_PFMG || type == OPF_SOLVER_SMG || type == OPF_SOLVER_JACOBI) ? OPF_SOLVER_PCG : type;
        int rc = run_iteration(s, ptype, w, bnorm, tol, maxit, &iters, &rel);
        s->params.precond = saved_pre;
        s->pin_active = false;
        return finish(rc);
    }
    if (s->pinned) {
        s->pin_active = true;
        poke(s, s->B, s->pin_off, 0.0);
        poke(s, s->X, s->pin_off, 0.0);
    }
    int rc = run_iteration(s, type, w, bnorm, tol, maxit, &iters, &rel);
    s->pin_active = false;
    return finish(rc);
}

}// extern "C"
