// engine.hpp -- internal structures of libopflow_b200.so (host side of the C ABI in include/opflow_b200.h).
#pragma once
#include "../../include/opflow_b200.h"
#include "opf_device.cuh"
#include <array>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace opfe {

    constexpr int D3 = OPF_MAX_DIM;

    // DS::Range<d> with stride 1 (src/DataStructures/Range/Ranges.hpp:33-225); unused axes are [0,1)
    struct Range {
        int start[D3] = {0, 0, 0}, end[D3] = {1, 1, 1};
        long long count() const {
            long long c = 1;
            for (int d = 0; d < D3; ++d) {
                if (end[d] - start[d] <= 0) return 0;
                c *= end[d] - start[d];
            }
            return c;
        }
        bool empty() const { return count() <= 0; }
        Range inner(int w, int dim) const {// getInnerRange(w) Ranges.hpp:170-178
            Range r = *this;
            for (int d = 0; d < dim; ++d) {
                r.start[d] += w;
                r.end[d] -= w;
            }
            return r;
        }
        bool operator==(const Range& o) const {
            for (int d = 0; d < D3; ++d)
                if (start[d] != o.start[d] || end[d] != o.end[d]) return false;
            return true;
        }
        void set_empty(int dim) {// Range::EmptyRange() Ranges.hpp:47-54
            for (int d = 0; d < dim; ++d) start[d] = end[d] = 0;
        }
        bool covers(const Range& o) const {
            for (int d = 0; d < D3; ++d)
                if (start[d] > o.start[d] || end[d] < o.end[d]) return false;
            return true;
        }
    };
    // DS::commonRange (Ranges.hpp:234-262): per-axis max(start), min(end) -- NOT clamped to non-negative extent
    inline Range common(const Range& a, const Range& b) {
        Range r;
        for (int d = 0; d < D3; ++d) {
            r.start[d] = a.start[d] > b.start[d] ? a.start[d] : b.start[d];
            r.end[d] = a.end[d] < b.end[d] ? a.end[d] : b.end[d];
        }
        return r;
    }
    inline opf_range to_c(const Range& r) {
        opf_range o;
        for (int d = 0; d < D3; ++d) {
            o.start[d] = r.start[d];
            o.end[d] = r.end[d];
        }
        return o;
    }
    inline Range from_c(const opf_range& r, int dim) {
        Range o;
        for (int d = 0; d < dim; ++d) {
            o.start[d] = r.start[d];
            o.end[d] = r.end[d];
        }
        return o;
    }

    // ---------------------------------------------------------------------------------------- context
    struct Context {
        bool inited = false;
        int device = -1;
        cudaStream_t stream = nullptr, comm_stream = nullptr;
        cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_comm = nullptr, ev_compute = nullptr;
        int mode = OPF_MODE_FAST;
        long long launches = 0;
        double* red_buf = nullptr;// partials + result
        int red_cap = 0;
        double* red_host = nullptr;// pinned
        int sm_count = 148;
        double* stage = nullptr;// dense staging box of opf_field_upload / download (grow-only)
        long long stage_elems = 0;
    };
    Context& ctx();
    int fail(int code, const char* fmt, ...);
    int require_device();
#define OPF_CUDA(call)                                                                                                 \
    do {                                                                                                               \
        cudaError_t e__ = (call);                                                                                      \
        if (e__ != cudaSuccess) return opfe::fail(OPF_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

    // ---------------------------------------------------------------------------------------- mesh
    struct AxisArrays {
        // host copies, index 0 <-> ext_range.start (CartesianMesh::_x/_dx/_idx, CartesianMesh.hpp:39-49)
        std::vector<double> x, dx, idx;
        std::vector<double> rdx, rdxh, rdxc;// Fast-mode reciprocals (opf_device.cuh AxisView)
        double* dev = nullptr;              // one allocation: x | dx | rdx | rdxh | rdxc, each n_ext long
        bool set = false;
        bool uniform = false;// every dx entry bitwise equal (opf::AxisView::uniform)
    };
}// namespace opfe

struct opf_mesh_s {
    int dim = 0;
    int dims[opfe::D3] = {1, 1, 1};
    int start[opfe::D3] = {0, 0, 0};
    int pad_width = 5;
    int ext_mode[opfe::D3] = {0, 0, 0};
    opfe::Range range, ext_range;
    opfe::AxisArrays ax[opfe::D3];
    int refcount = 1;
    bool device_ready = false;
};

namespace opfe {
    int mesh_upload(opf_mesh_s* m);
    opf::AxisView mesh_axis_view(const opf_mesh_s* m, int d);

    // ---------------------------------------------------------------------------------------- field
    struct BC {
        int type = OPF_BC_UNDEFINED;
        double value = 0;
        std::vector<double> face;// functor BC values over face_range
        Range face_range;
        double* face_dev = nullptr;
    };
    struct Neighbor {// internal::NeighborInfo (StructuredFieldExpr.hpp:28-36)
        int rank;
        Range send, recv;
        int code;
    };
    // one ghost-fill launch: see fill_kernel in engine_core.cu
    struct FillOp {
        int kind; // 0 set-bc, 1 dirc, 2 neum, 3 symm, 4 asymm, 5 periodic copy
        int axis, side, center;
        Range r;
        int mirror_c;// mirror index = mirror_c - idx[axis]   (periodic: shift = mirror_c)
        int xb;      // boundary node index used by the Dirichlet mid-point rule
        const BC* bc;
    };
}// namespace opfe

struct opf_field_s {
    std::string name;
    int dim = 0;
    opf_mesh_s* mesh = nullptr;
    int loc[opfe::D3] = {0, 0, 0};
    opfe::BC bc[opfe::D3][2];
    int ext[opfe::D3][2] = {{0, 0}, {0, 0}, {0, 0}};
    int padding = 0;
    opfe::Range local, assignable, accessible, logical, storage;
    int n_ranks = 1, rank = 0;
    std::vector<opfe::Range> split_map; // per rank, Corner fields already carry the extra end node
    std::vector<opfe::Range> cell_split;// the strategy's own cell-centred blocks (AbstractSplitStrategy::getSplitMap), empty on one rank
    std::vector<opfe::Neighbor> neighbors;
    // storage: element with global index g lives at buf[cur][lead + (g0-S0) + (g1-S1)*pitch1 + (g2-S2)*pitch2]
    double* buf[2] = {nullptr, nullptr};
    int cur = 0;
    long long pitch1 = 0, pitch2 = 0, lead = 0, elems = 0;
    // Guard doubles allocated (and zeroed) before buf[i] and after buf[i] + elems.  The register-window skeleton stages the UNION of
    // the rows / planes an operator may tap over all staggerings (the location is a run-time property, the tap set a compile-time
    // one), up to 3 rows and 3 planes beyond the storage range; those values are never consumed, but the loads must land in mapped
    // memory.  buf[i] itself is what every other piece of code sees (128-byte aligned).
    long long guard = 0;
    std::vector<opfe::FillOp> fill0, fill1, fill2;// step 0, step 1, step 2 (single-rank periodic)
    // step 0 writes time-independent values (ConstDircBC / pre-evaluated FunctorDircBC) into Corner boundary nodes that no
    // assignment kernel ever touches (they lie outside assignableRange): once a buffer holds them they stay valid until
    // someone writes the buffer from outside (upload, swap, BC change), so the launch is skipped while this flag is set.
    bool bc0_clean[2] = {false, false};
    // halo staging (multi-rank)
    double* halo_send = nullptr;
    double* halo_recv = nullptr;
    long long halo_elems = 0;

    double* biased(int which) const {// pointer such that p[g0 + g1*pitch1 + g2*pitch2] is element g
        return buf[which] + lead - ((long long) storage.start[0] + (long long) storage.start[1] * pitch1 + (long long) storage.start[2] * pitch2);
    }
    double* first() const { return buf[cur] + lead; }
};

namespace opfe {
    int field_update_padding(opf_field_s* f);
    int field_fill_bc(opf_field_s* f, const Range* clip, cudaStream_t st = nullptr);// steps 0-1 of updatePadding, optionally clipped to a box
    int field_fill_periodic(opf_field_s* f);// step 2 local part: periodic copies of the axes that are not split across ranks
    int field_ensure_twin(opf_field_s* f);
    // f takes over g's decomposition, ranges and device storage; g is left without storage (the caller destroys it)
    void field_adopt(opf_field_s* f, opf_field_s* g);
    void repeat_cache_clear();// engine_expr.cu: captured opf_assign_repeat graphs hold buffer addresses
    // dense box (strides 1, d1, d2) <-> pitched storage of buffer `which` over range r, on stream st
    int dense_convert(opf_field_s* f, int which, double* dense, const Range& r, long long d1, long long d2, bool unpack, cudaStream_t st);
    int halo_exchange(opf_field_s* f, cudaStream_t st);// engine_comm.cu: pack, NCCL send/recv group, unpack -- all on `st`
    void compute_neighbors(opf_field_s* f);
    bool comm_active();
    int comm_allreduce_device(double* dev, int n, int rop, cudaStream_t st);// in place, stream-ordered, no host synchronisation
}// namespace opfe
