/* opflow_b200.h -- C ABI of the B200-native evaluation engine for OpFlow's stencil hot path.
 *
 * The reference (OpFlow v0.2.7, header-only C++) has no binary plugin ABI; its drop-in boundary is a set of
 * compile-time customisation points.  This header is the C ABI that sits *underneath* those
 * customisation points: the templated front-end headers under opflow_b200/include/ (`#include <OpFlow>`) call only
 * these entry points, and so do the ctypes bindings used by tests/ and bench.py.  Every entry point cites the
 * reference interface it replaces (paths relative to the reference tree).
 *
 * Conventions
 *   - all handles are opaque pointers owned by the library; *_destroy releases them.
 *   - every function that can fail returns an int status (OPF_OK == 0) unless it returns a handle (NULL on
 *     failure) -- opf_last_error() gives the message.  There is NO CPU fallback: without a CUDA device every
 *     compute entry point fails with OPF_ERR_NO_DEVICE.
 *   - ranges are half-open [start, end) per axis, axis 0 fastest, exactly DS::Range<d>
 *     (src/DataStructures/Range/Ranges.hpp:33-225); unused axes have start=0,end=1.
 *   - index arithmetic visible through this ABI is 32-bit `int` like the reference (src/Core/BasicDataTypes.hpp:23).
 */
#ifndef OPFLOW_B200_H
#define OPFLOW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OPF_MAX_DIM 3
#define OPF_MAX_FIELDS 32  /* field leaves per expression  */
#define OPF_MAX_SCALARS 32 /* scalar leaves per expression (a 3x3x3 convolution kernel takes 27) */
#define OPF_MAX_NODES 96   /* tree nodes per expression    */

typedef enum {
    OPF_OK = 0,
    OPF_ERR_NO_DEVICE = 1,   /* no CUDA device / driver: the engine never falls back to the CPU */
    OPF_ERR_INVALID = 2,     /* bad argument */
    OPF_ERR_UNSUPPORTED = 3, /* expression signature not compiled in (see opf_expr_register) */
    OPF_ERR_RANGE = 4,       /* expression would read outside a leaf field's storage */
    OPF_ERR_CUDA = 5,
    OPF_ERR_LOC = 6,         /* operands' LocOnMesh differ (reference aborts: BinOpDefMacros.hpp.in:25-42) */
    OPF_ERR_COMM = 7,
    OPF_ERR_NOT_CONVERGED = 8
} opf_status;

/* src/Core/Constants.hpp:48-53, src/Core/BC/BCBase.hpp:24 -- numeric values identical to the reference enums */
typedef enum { OPF_LOC_CORNER = 0, OPF_LOC_CENTER = 1 } opf_loc;
typedef enum { OPF_POS_START = 0, OPF_POS_END = 1 } opf_dimpos;
typedef enum { OPF_BC_UNDEFINED = 0, OPF_BC_DIRC = 1, OPF_BC_NEUM = 2, OPF_BC_PERIODIC = 3, OPF_BC_INTERNAL = 4,
               OPF_BC_SYMM = 5, OPF_BC_ASYMM = 6 } opf_bctype;
typedef enum { OPF_OP_EQ = 0, OPF_OP_ADD = 1, OPF_OP_MINUS = 2, OPF_OP_MUL = 3, OPF_OP_DIV = 4 } opf_assign_op;
typedef enum { OPF_MESHEXT_UNDEFINED = 0, OPF_MESHEXT_SYMM = 1, OPF_MESHEXT_PERIODIC = 2, OPF_MESHEXT_UNIFORM = 3 } opf_meshext;
/* arithmetic policy of the device functors (DESIGN.md "exact vs fast"):
 *   EXACT  every +,-,*,/ is the IEEE-rn operation in the reference's order, no FMA contraction: explicit updates are
 *          bit-identical to the reference's CPU path.
 *   FAST   per-axis reciprocal coefficient arrays replace the divides of D2SecondOrderCentered / WENO53; FMA allowed;
 *          within 1e-12 relative of the reference (BASELINE.json north_star tolerance).
 *   STENCIL the reference's implicit-path arithmetic: EXACT, except that a / b is a * (1. / b) as in StencilPad::operator/
 *          (StencilPad.hpp:293-296).  Probing the matrix-free operator in this mode yields the reference's assembled CSR / HYPRE
 *          coefficients and right-hand side bit for bit on any mesh; EXACT reproduces them only where the reciprocals are exact
 *          (dx a power of two).  Verification mode: direct-global skeleton only. */
typedef enum { OPF_MODE_EXACT = 0, OPF_MODE_FAST = 1, OPF_MODE_STENCIL = 2 } opf_mode;
typedef enum { OPF_RED_SUM = 0, OPF_RED_MAX = 1, OPF_RED_MIN = 2, OPF_RED_ABSMAX = 3, OPF_RED_SUMSQ = 4 } opf_reduce_op;

typedef struct opf_range { int start[OPF_MAX_DIM], end[OPF_MAX_DIM]; } opf_range;

typedef struct opf_mesh_s* opf_mesh_t;
typedef struct opf_field_s* opf_field_t;
typedef struct opf_solver_s* opf_solver_t;

/* ------------------------------------------------------------------------------------------------ runtime */
/* Environment.hpp:29-62 (InitEnvironment / global ParallelPlan).  device < 0: use LOCAL_RANK or 0. */
int opf_init(int device);
int opf_finalize(void);
const char* opf_last_error(void);
const char* opf_version(void);
int opf_device_count(void);
int opf_set_mode(int mode);         /* opf_mode; default OPF_MODE_FAST */
int opf_get_mode(void);
int opf_synchronize(void);
void* opf_stream(void);             /* the engine's compute cudaStream_t (for callers that launch their own kernels) */
long long opf_launch_count(void);   /* kernels launched by this library since opf_init (bench.py gpu_launches) */
/* Name of the kernel skeleton the most recent assignment launched ("opf::tma_kernel", "opf::tma2d_kernel", "opf::window_kernel",
 * "opf::assign_kernel"): lets bench.py and the tests report what actually ran instead of assuming it. */
const char* opf_last_kernel_name(void);
/* Run-time switches for A/B runs and tests; each starts from the environment variable OPF_<KEY> or its default (1):
 * "tma", "tma2d", "window" (skeleton selection), "overlap" (halo exchange overlapped with the interior sweep), "graphs" (CUDA-graph
 * replay of the multigrid cycle), "mg_fused", "direct_halo" (contiguous slab faces sent straight from field storage),
 * "fused_krylov" (device-resident PCG scalars).  opf_get_option returns -1 for an unknown key. */
int opf_set_option(const char* key, int value);
int opf_get_option(const char* key);
/* CUDA-event timing on the engine's compute stream (bench.py): returns elapsed ms between begin and end */
int opf_timer_begin(void);
int opf_timer_end(float* ms);

/* ------------------------------------------------------------------------------------------------ mesh */
/* MeshBuilder<CartesianMesh<Dim>>: newMesh (CartesianMesh.hpp:132-139), setStart :155, setPadWidth :150,
 * setExtMode :178-186, setMeshOfDim(k,min,max) :168-171 -> set1DMesh :292-303, setMeshOfDim(k,f) :163-166
 * -> set1DMesh :279-290 (here the node coordinates f(i) are passed pre-evaluated). */
opf_mesh_t opf_mesh_create(int dim, const int* dims, const int* start, int pad_width);
int opf_mesh_set_ext_mode(opf_mesh_t m, int axis, int mode);
int opf_mesh_set_uniform(opf_mesh_t m, int axis, double xmin, double xmax);
int opf_mesh_set_coords(opf_mesh_t m, int axis, const double* x, int n);
int opf_mesh_get_range(opf_mesh_t m, opf_range* range, opf_range* ext_range);
/* copies x (ext length n_ext) / dx, idx (n_ext-1) of one axis; returns the number of doubles written */
int opf_mesh_get_axis(opf_mesh_t m, int axis, double* x, double* dx, double* idx, int cap);
int opf_mesh_destroy(opf_mesh_t m);

/* ------------------------------------------------------------------------------------------------ field */
typedef struct opf_bc_desc {
    int type;            /* opf_bctype */
    double value;        /* ConstDircBC / ConstNeumBC value (DircBC.hpp:40-50) */
    const double* face;  /* FunctorDircBC / FunctorNeumBC pre-evaluated on the host (DircBC.hpp:83-118): one value per
                            index of the face slab `face_range`, axis 0 fastest; NULL for const / logical BCs */
    opf_range face_range;
} opf_bc_desc;

typedef struct opf_field_desc {
    opf_mesh_t mesh;
    int loc[OPF_MAX_DIM];                 /* opf_loc per axis  (ExprBuilder::setLoc, CartesianField.hpp:811-824) */
    opf_bc_desc bc[OPF_MAX_DIM][2];       /* [axis][opf_dimpos] (setBC :827-895) */
    int ext[OPF_MAX_DIM][2];              /* setExt :897-913 */
    int padding;                          /* setPadding :915-918 */
    /* decomposition (setSplitStrategy :920-923).  n_ranks<=1: no strategy (localRange = accessibleRange).
     * Otherwise split_map holds the strategy's getSplitMap() result (cell-centred ranges, one per rank:
     * AbstractSplitStrategy.hpp:24-32) -- build it with opf_split_even / opf_split_slab or by hand
     * (ManualSplitStrategy.hpp:34-57). */
    int n_ranks, rank;
    const opf_range* split_map;
} opf_field_desc;

/* ExprBuilder<CartesianField>::build (CartesianField.hpp:929-938): calculateRanges :950-1029, validateRanges
 * :941-948, storage = localRange inflated by padding, updatePadding(). */
opf_field_t opf_field_create(const opf_field_desc* desc, const char* name);
/* the host half of build(): ranges, split and neighbour lists of a description WITHOUT device storage (works with no GPU; the
 * handle answers opf_field_get_range / get_loc / padding / neighbors / destroy, every compute entry point rejects it). */
opf_field_t opf_field_plan(const opf_field_desc* desc, const char* name);
opf_field_t opf_field_clone(opf_field_t f, const char* name); /* CartesianField copy ctor :57-68 (deep copy) */
int opf_field_destroy(opf_field_t f);
int opf_field_dim(opf_field_t f);
/* which: 0 local, 1 assignable, 2 accessible, 3 logical, 4 storage (local inflated by padding), 5 local-readable */
int opf_field_get_range(opf_field_t f, int which, opf_range* out);
int opf_field_get_loc(opf_field_t f, int* loc);
int opf_field_padding(opf_field_t f);
/* device storage: pointer to the element with global index (0,0,0) may lie outside the allocation, so the ABI exposes
 * the pointer to the first stored element (index = storage.start) plus pitches in elements (pitch0 == 1). */
int opf_field_device_ptr(opf_field_t f, double** first, long long* pitch1, long long* pitch2);
/* host <-> device transfer of the values over `range` (axis 0 fastest, dense).  PlainTensor is the host twin
 * (src/DataStructures/Arrays/Tensor/PlainTensor.hpp:204-243). */
int opf_field_upload(opf_field_t f, const opf_range* range, const double* host);
int opf_field_download(opf_field_t f, const opf_range* range, double* host);
/* assignImpl_final(const D&) CartesianField.hpp:237-280: field (op)= c over assignable∩local, then updatePadding */
int opf_field_assign_scalar(opf_field_t f, int op, double c);
/* Asynchronous snapshot for writers (reference: the `<<` of src/Utils/Writers/{RawBinaryStream,HDF5Stream,TecplotASCIIStream}.hpp reads
 * the field on the host; here the values of `range` -- default localRange -- as of this point of the program are packed on the
 * device and copied to `pinned_host` (opf_host_alloc) on a copy stream while the caller goes on; dense, axis 0 fastest).
 * opf_snapshot_wait blocks until the host buffer is complete and releases the handle. */
typedef struct opf_snapshot_s* opf_snapshot_t;
void* opf_host_alloc(unsigned long long bytes); /* pinned host memory */
int opf_host_free(void* p);
opf_snapshot_t opf_field_snapshot(opf_field_t f, const opf_range* range, double* pinned_host);
int opf_snapshot_wait(opf_snapshot_t s);
/* assignImpl_final(const CartesianField&) :180-193: dst (op)= src (both initialised) */
int opf_field_assign_field(opf_field_t dst, int op, opf_field_t src);
/* updatePaddingImpl_final CartesianField.hpp:349-769: step 0 corner-Dirichlet nodes, step 1 BC ghost extension,
 * step 2 periodic copy (single rank) or halo exchange (multi rank, NCCL). */
int opf_field_update_padding(opf_field_t f);
/* replace a const BC value (keeps type) -- used by tests that mutate BCs */
int opf_field_set_bc_value(opf_field_t f, int axis, int pos, double value);
/* std::swap(CartesianField&, CartesianField&) CartesianField.hpp:1039-1041: swaps storage only */
int opf_field_swap(opf_field_t a, opf_field_t b);
/* CartesianField::resplitWithStrategy (CartesianField.hpp:83-177): move a decomposed field, values kept, to the decomposition given by
 * split_map (n_ranks cell-centred blocks like opf_field_desc.split_map).  Collective over the communicator; the handle stays valid, its
 * ranges / neighbours / storage are those of a field built with the new map.  Solvers created on the field must be re-created.  Fields
 * with functor boundary values are refused; a non-decomposed field is left alone (the reference's method only acts under MPI). */
int opf_field_resplit(opf_field_t f, const opf_range* split_map);
/* the host half of opf_field_resplit, usable on a plan (no device): for this rank, the box it sends to and the box it receives from every
 * rank r (send[r], recv[r]: n_ranks entries each, empty boxes where nothing moves; entry [rank] is the part that stays) and its
 * localRange under the new map. */
int opf_field_resplit_plan(opf_field_t f, const opf_range* split_map, opf_range* send, opf_range* recv, opf_range* new_local);
/* number of neighbours and their (rank, send, recv, shift-code) tuples: updateNeighbors :298-347 */
int opf_field_neighbors(opf_field_t f, int cap, int* ranks, opf_range* send, opf_range* recv, int* codes);

/* ------------------------------------------------------------------------------------------------ expressions */
/* An expression is named by its *signature*: the C++ type of its device functor, e.g. FTCS2D
 * (examples/FTCS2D/FTCS-OMP.cpp:26)   Add<F<0>,Mul<S<0>,Add<D2C<0,F<1>>,D2C<1,F<2>>>>>
 * Grammar (opflow_b200/csrc/opf_device.cuh): leaves F<k> (k-th field argument), S<k> (k-th scalar argument);
 * point-wise Add Sub Mul Div Min Max Pow Lt Le Gt Ge Eq Ne And Or <A,B>, Neg Pos Not Sqrt Abs Exp Log Sin Cos Tan
 * Tanh Pow2 <A>, Cond<C,A,B>; stencils D2C<d,E> (D2SecondOrderCentered), D1C<d,E> (D1FirstOrderCentered),
 * D1Dn<d,E>/D1Up<d,E> (D1FirstOrderBiasedDownwind/Upwind), WenoDn<d,E>/WenoUp<d,E> (D1WENO53Downwind/Upwind),
 * IntpC2N<d,E>/IntpN2C<d,E> (D1Linear Cen2Cor / Cor2Cen).
 * The front-end headers instantiate the kernels for their expression types with nvcc and register them here; the
 * library itself carries the expressions of the acceptance programs (opf_expr_builtin_count/name). */
typedef int (*opf_expr_launcher)(const void* args_blob, const void* launch_blob, void* stream);
int opf_expr_register(const char* signature, opf_expr_launcher fn);
/* same, with the layout stamp of the translation unit that instantiated the launcher (OPF_DEVICE_ABI of opf_device.cuh: sizes of the
 * ExprArgs / LaunchInfo blobs the engine hands to it): a program compiled against an older header is refused with OPF_ERR_INVALID
 * instead of reading the blobs with the wrong layout.  The front-end headers call this one. */
int opf_expr_register_abi(const char* signature, opf_expr_launcher fn, unsigned long long abi);
int opf_expr_is_registered(const char* signature);
int opf_expr_builtin_count(void);
const char* opf_expr_builtin_name(int i);

/* Expr::prepare() range algebra (Expression.hpp:99-103 + every Op::prepare, e.g. D2SecondOrderCentered.hpp:187-204,
 * BinOpDefMacros.hpp.in:19-59): ranges/loc of the prepared expression.  which: 0 local, 2 accessible, 3 logical. */
int opf_expr_prepare(const char* signature, const opf_field_t* fields, int nfields, int which, opf_range* out,
                     int* loc);
/* FieldAssigner::assign<Op>(src, dst) FieldAssigner.hpp:26-36 + assign_impl :40-86, followed by dst.updatePadding()
 * as in CartesianField::assignImpl_final :195-234.  Aliasing (src.contains(dst)) behaves as if the RHS were fully
 * evaluated first (the engine writes the twin buffer and swaps instead of copying). */
int opf_assign(opf_field_t dst, int op, const char* signature, const opf_field_t* fields, int nfields,
               const double* scalars, int nscalars);
/* same, with flags: OPF_ASSIGN_NO_PADDING skips the trailing updatePadding() (solver work vectors whose ghosts are
 * refreshed explicitly before the next operator application) */
#define OPF_ASSIGN_NO_PADDING 1
int opf_assign_ex(opf_field_t dst, int op, const char* signature, const opf_field_t* fields, int nfields,
                  const double* scalars, int nscalars, int flags);
/* `count` consecutive identical assignments -- a time loop whose body is this one statement (examples/FTCS2D/FTCS-OMP.cpp:24-27) --
 * replayed from a CUDA graph of 32 steps each (captured once per destination / expression / operands / mode): removes the launch
 * gaps that bound small fields (BASELINE config C1: 1025^2).  Results are identical to calling opf_assign `count` times. */
int opf_assign_repeat(opf_field_t dst, int op, const char* signature, const opf_field_t* fields, int nfields, const double* scalars,
                      int nscalars, int count);
/* The same assignment for callers whose data lives in HOST memory (pinned for full speed): the values of `in_field` over its
 * localRange come from host_in (axis 0 fastest, dense), dst (op)= expr is evaluated, and dst's localRange is written to host_out.
 * in_field is dst itself or one of the expression's leaves; its ghost cells (BC extension, periodic images, halo planes of a
 * decomposed field) are refreshed from the uploaded values before the sweep, as updatePadding() would.  Upload, sweep and download
 * are pipelined in slabs along the slowest axis on three streams -- also for slab-decomposed fields, whose two boundary chunks travel
 * first so that the input's halo exchange starts early; returns when host_out is complete.  (The reference has no such call: its fields ARE host memory --
 * this is what a caller that keeps PlainTensor storage on the host, CartesianField.hpp:37, would use per assignment.) */
int opf_assign_host(opf_field_t dst, int op, const char* signature, const opf_field_t* fields, int nfields, const double* scalars,
                    int nscalars, opf_field_t in_field, const double* host_in, double* host_out);
/* rangeReduce(range, op, expr.evalAt) RangeFor.hpp:87-121 for a device expression; range==NULL: the expression's
 * local ∩ accessible range.  globalReduce (:125-135) = this + opf_comm_allreduce. */
int opf_reduce(int rop, const char* signature, const opf_field_t* fields, int nfields, const double* scalars,
               int nscalars, const opf_range* range, double* result);

/* ------------------------------------------------------------------------------------------------ decomposition */
/* EvenSplitStrategy<F>::getSplitMap (EvenSplitStrategy.hpp:57-192): mesh_range is the *nodal* mesh range; out gets
 * n_ranks cell-centred block ranges in rank order. */
int opf_split_even(int dim, const opf_range* mesh_range, int n_ranks, opf_range* out);
/* slab split along the slowest axis behind the same interface */
int opf_split_slab(int dim, const opf_range* mesh_range, int n_ranks, opf_range* out);

/* communicator: one process per GPU.  id is an ncclUniqueId (128 bytes) produced by opf_comm_unique_id on rank 0 and
 * broadcast by the caller (torch.distributed / MPI / files).  Replaces MPI_Isend/Irecv/Waitall of
 * CartesianField.hpp:689-730 and MPI_Allgather of RangeFor.hpp:132. */
int opf_comm_unique_id(void* id128);
int opf_comm_init(int rank, int n_ranks, const void* id128);
int opf_comm_rank(void);
int opf_comm_size(void);
int opf_comm_allreduce(double* values, int n, int rop);
int opf_comm_finalize(void);

/* ------------------------------------------------------------------------------------------------ implicit */
/* StructSolverType (src/Core/Solvers/Struct/StructSolver.hpp:23-34) -- numeric values identical */
typedef enum { OPF_SOLVER_NONE = 0, OPF_SOLVER_JACOBI = 1, OPF_SOLVER_SMG = 2, OPF_SOLVER_PFMG = 3, OPF_SOLVER_CYCRED = 4,
               OPF_SOLVER_PCG = 5, OPF_SOLVER_GMRES = 6, OPF_SOLVER_FGMRES = 7, OPF_SOLVER_LGMRES = 8,
               OPF_SOLVER_BICGSTAB = 9 } opf_solver_type;

typedef struct opf_solver_params { /* StructSolverParamsBase :39-51 + the per-solver fields the engine honours */
    int type, precond;      /* opf_solver_type */
    double tol;             /* relative residual ||r||2/||b||2 */
    int max_iter;
    int static_mat, pin_value;
    double precond_tol;     /* PFMG-as-preconditioner tolerance (StructSolverPFMG.hpp:23-34); 0 = one V-cycle */
    int precond_max_iter;
    int num_pre_relax, num_post_relax, relax_type; /* StructSolverPFMG.hpp:23-34 relaxType.  0 / 1: weighted Jacobi (omega = 2d/(2d+1));
                                                       2: symmetric red-black Gauss-Seidel (red-black before, black-red after the
                                                       coarse correction); 3: red-black on both sides */
    int print_level;
    int k_dim;              /* GMRES restart length (StructSolverGMRES.hpp kDim; 0 = hypre's default 5) */
} opf_solver_params;

typedef struct opf_solve_state { int niter; double relerr, abserr; } opf_solve_state; /* EqnSolveState EqnSolveHandler.hpp:17-25 */

/* makeEqnSolveHandler(f, target, solver) -> HYPREEqnSolveHandler ctor + init() (HYPREEqnSolveHandler.hpp:44-117), matrix-free:
 * the equation  lhs(e) == rhs  is given as two expression signatures.  `lhs` must be linear in the unknown e; the leaves of
 * lhs that ARE the unknown are flagged in `unknown_mask` (bit k <=> field leaf k is e; their entries in lhs_fields are ignored).
 * The operator is never assembled: A.p = lhs evaluated on p with the target's boundary conditions made homogeneous, and
 * b = rhs - lhs(e = 0 with the real boundary data).  pin_value pins the first assignable cell exactly like the reference
 * (HYPREEqnSolveHandler.hpp:145-163, StencilField.hpp:132).  Supported: type PCG / BICGSTAB / GMRES(k) (FGMRES and LGMRES requests run as GMRES) / JACOBI / PFMG
 * (stand-alone geometric multigrid), precond NONE / JACOBI / PFMG.  lhs may also carry terms without the unknown (affine operator:
 * the front-end passes lhs(e) - rhs(e) with rhs "S<0>" = 0 when both sides of `==` contain e).  Multigrid on an lhs with coefficient
 * fields restricts those fields level by level (one rank; constant boundary values; used when the level-0 diagonal does not dominate),
 * otherwise such a request runs its level-0 smoother.  A decomposed target (opf_field_desc.split_map) with an lhs whose field leaves
 * are all the unknown is solved with distributed multigrid levels. */
opf_solver_t opf_solver_create(opf_field_t target, const char* lhs_signature, const opf_field_t* lhs_fields, int n_lhs_fields,
                               const double* lhs_scalars, int n_lhs_scalars, unsigned unknown_mask,
                               const opf_solver_params* params);
/* EqnSolveHandler::solve() (HYPREEqnSolveHandler.hpp:190-209): evaluates rhs, solves, writes the solution into the target field
 * and refreshes its padding (returnValues :181-188).  Initial guess = current target values (initx :119-123). */
int opf_solver_solve(opf_solver_t s, const char* rhs_signature, const opf_field_t* rhs_fields, int n_rhs_fields,
                     const double* rhs_scalars, int n_rhs_scalars, opf_solve_state* state);
/* refresh the non-unknown leaves of lhs (coefficient fields, scalars such as dt) before the next solve: the handler's equation
 * lambda captures them by reference and is re-evaluated on every solve() (HYPREEqnSolveHandler.hpp:190-209 -> generateAb). */
int opf_solver_update(opf_solver_t s, const opf_field_t* lhs_fields, int n_lhs_fields, const double* lhs_scalars, int n_lhs_scalars);
int opf_solver_levels(opf_solver_t s); /* number of multigrid levels built (1 = no hierarchy) */
/* CSRMatrixGenerator::generate (src/Core/Equation/CSRMatrixGenerator.hpp:55-138) for callers that want the assembled system (an
 * algebraic solver, a dump): row pointers [rows + 1], ascending column indices and values [cap_nnz], right-hand side [rows], rows =
 * assignable cells in x-fastest order; pin_last replaces the last row by the identity like the reference's CSR route.  The values are
 * probed off the matrix-free operator; in OPF_MODE_STENCIL they are the reference's assembled coefficients bit for bit. */
int opf_solver_export_csr(opf_solver_t s, const char* rhs_signature, const opf_field_t* rhs_fields, int n_rhs_fields, const double* rhs_scalars,
                          int n_rhs_scalars, int pin_last, long long cap_nnz, int* ptr, int* col, double* val, double* rhs, long long* nnz_out);
int opf_solver_destroy(opf_solver_t s);

#ifdef __cplusplus
}
#endif
#endif /* OPFLOW_B200_H */
