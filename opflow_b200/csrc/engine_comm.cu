// engine_comm.cu -- one-process-per-GPU communicator (NCCL over NVLink5/NVSwitch) and the halo exchange that replaces
// the MPI branch of updatePaddingImpl_final (CartesianField.hpp:630-768: serial pack -> MPI_Isend/Irecv -> Waitall ->
// serial unpack) and the MPI_Allgather of globalReduce (RangeFor.hpp:125-135).
// NCCL is dlopen'ed so single-GPU use has no NCCL dependency.
#include "engine.hpp"
#include <algorithm>
#include <dlfcn.h>

namespace {
    typedef struct ncclComm* ncclComm_t;
    typedef struct { char internal[128]; } ncclUniqueId;
    typedef int ncclResult_t;
    enum { ncclFloat64 = 8 };
    enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 };

    struct Nccl {
        void* h = nullptr;
        ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
        ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
        ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
        ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
        ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
        ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
        ncclResult_t (*GroupStart)() = nullptr;
        ncclResult_t (*GroupEnd)() = nullptr;
        const char* (*GetErrorString)(ncclResult_t) = nullptr;
        ncclComm_t comm = nullptr;
        int rank = 0, size = 1;
        double* dbuf = nullptr;
    };
    Nccl& nc() {
        static Nccl n;
        return n;
    }
    int load_nccl() {
        Nccl& n = nc();
        if (n.h) return OPF_OK;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            n.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (n.h) break;
        }
        if (!n.h) return opfe::fail(OPF_ERR_COMM, "cannot dlopen libnccl.so.2: %s", dlerror());
#define OPF_SYM(field, sym)                                                                                            \
    *(void**) (&n.field) = dlsym(n.h, sym);                                                                            \
    if (!n.field) return opfe::fail(OPF_ERR_COMM, "libnccl lacks %s", sym);
        OPF_SYM(GetUniqueId, "ncclGetUniqueId")
        OPF_SYM(CommInitRank, "ncclCommInitRank")
        OPF_SYM(CommDestroy, "ncclCommDestroy")
        OPF_SYM(Send, "ncclSend")
        OPF_SYM(Recv, "ncclRecv")
        OPF_SYM(AllReduce, "ncclAllReduce")
        OPF_SYM(GroupStart, "ncclGroupStart")
        OPF_SYM(GroupEnd, "ncclGroupEnd")
        OPF_SYM(GetErrorString, "ncclGetErrorString")
#undef OPF_SYM
        return OPF_OK;
    }
#define OPF_NCCL(call)                                                                                                 \
    do {                                                                                                               \
        ncclResult_t r__ = (call);                                                                                     \
        if (r__ != 0) return opfe::fail(OPF_ERR_COMM, "%s: %s", #call, nc().GetErrorString(r__));                     \
    } while (0)

    int inverse_code(int code, int dim) {
        int out = 0, p = 1;
        for (int d = 0; d < dim; ++d) {
            int dir = (code / p) % 3;
            if (dir == 1) dir = 2;
            else if (dir == 2)
                dir = 1;
            out += dir * p;
            p *= 3;
        }
        return out;
    }
}// namespace

namespace opfe {
    bool comm_active() { return nc().comm != nullptr; }

    int comm_allreduce_device(double* dev, int cnt, int rop, cudaStream_t st) {
        Nccl& n = nc();
        if (!n.comm || cnt <= 0) return OPF_OK;
        const int op = rop == OPF_RED_MAX || rop == OPF_RED_ABSMAX ? ncclMax : (rop == OPF_RED_MIN ? ncclMin : ncclSum);
        OPF_NCCL(n.AllReduce(dev, dev, (size_t) cnt, ncclFloat64, op, n.comm, st));
        return OPF_OK;
    }

    int halo_exchange(opf_field_s* f, cudaStream_t st) {
        Nccl& n = nc();
        if (f->neighbors.empty()) return OPF_OK;
        if (!n.comm) return fail(OPF_ERR_COMM, "field '%s' is decomposed over %d ranks but opf_comm_init was not called", f->name.c_str(), f->n_ranks);
        const int nn = (int) f->neighbors.size();
        // Direct faces.  When the decomposition only cuts the slowest axis (slabs), every rank's storage has the same pitches and
        // lead, and the `padding` planes (rows in 2-D) next to a neighbour are ONE contiguous run of the pitched buffer: they are
        // sent straight from the field and received straight into the ghost planes -- no pack / unpack kernels, no staging
        // ("each face is one contiguous block").  The run carries the planes' own ghost cells of the unsplit axes too;
        // they land in the receiver's corner ghosts, which the reference leaves untouched and no stencil ever reads before the
        // next fill.  Anything else (blocks, pencils, periodic images of a split axis with odd shapes) takes the staged route.
        const int ax = f->dim - 1;
        bool slab_layout = opf_internal_opt(OPF_OPT_DIRECT_HALO) != 0 && f->dim >= 2;
        for (const auto& r : f->split_map)
            for (int d = 0; d < ax && slab_layout; ++d)
                if (r.start[d] != f->split_map[0].start[d] || r.end[d] != f->split_map[0].end[d]) slab_layout = false;
        const long long plane = f->dim == 3 ? f->pitch2 : f->pitch1;// doubles per index of the slowest axis
        auto spans_lower_axes = [&](const Range& r) {
            for (int d = 0; d < ax; ++d)
                if (r.start[d] != f->local.start[d] || r.end[d] != f->local.end[d]) return false;
            return true;
        };
        std::vector<char> direct(nn, 0);
        for (int i = 0; i < nn; ++i) {
            const auto& nb = f->neighbors[i];
            direct[i] = slab_layout && spans_lower_axes(nb.send) && nb.recv.count() > 0 && spans_lower_axes(nb.recv)
                        && nb.send.end[ax] - nb.send.start[ax] == nb.recv.end[ax] - nb.recv.start[ax]
                        && nb.recv.start[ax] >= f->storage.start[ax] && nb.recv.end[ax] <= f->storage.end[ax];
        }
        auto run_ptr = [&](int k) { return f->buf[f->cur] + f->lead + (long long) (k - f->storage.start[ax]) * plane; };
        // staging layout: sends in neighbour order, recvs in neighbour order
        std::vector<long long> soff(nn + 1, 0), roff(nn + 1, 0);
        for (int i = 0; i < nn; ++i) {
            soff[i + 1] = soff[i] + (direct[i] ? 0 : f->neighbors[i].send.count());
            roff[i + 1] = roff[i] + (direct[i] ? 0 : std::max<long long>(0, f->neighbors[i].recv.count()));
        }
        const long long need = std::max(soff[nn], roff[nn]);
        if (need > f->halo_elems) {
            if (f->halo_send) cudaFree(f->halo_send);
            if (f->halo_recv) cudaFree(f->halo_recv);
            OPF_CUDA(cudaMalloc(&f->halo_send, sizeof(double) * need));
            OPF_CUDA(cudaMalloc(&f->halo_recv, sizeof(double) * need));
            f->halo_elems = need;
        }
        auto dense = [&](const Range& r, double* stage, bool unpack) {// rows / planes from the grid, axis 0 coalesced
            const long long n0 = r.end[0] - r.start[0], n1 = r.end[1] - r.start[1];
            return dense_convert(f, f->cur, stage, r, n0, n0 * n1, unpack, st);
        };
        for (int i = 0; i < nn; ++i)
            if (!direct[i])
                if (int rc = dense(f->neighbors[i].send, f->halo_send + soff[i], false)) return rc;
        // message order per peer: sender's shift code ascending on both sides (the reference matches by tag =
        // hash(recv range), CartesianField.hpp:689-716)
        std::vector<int> sorder(nn), rorder(nn);
        for (int i = 0; i < nn; ++i) sorder[i] = rorder[i] = i;
        std::stable_sort(sorder.begin(), sorder.end(), [&](int a, int b) { return f->neighbors[a].code < f->neighbors[b].code; });
        std::stable_sort(rorder.begin(), rorder.end(), [&](int a, int b) {
            return inverse_code(f->neighbors[a].code, f->dim) < inverse_code(f->neighbors[b].code, f->dim);
        });
        OPF_NCCL(n.GroupStart());
        for (int k = 0; k < nn; ++k) {
            const int i = sorder[k];
            const auto& nb = f->neighbors[i];
            if (direct[i]) OPF_NCCL(n.Send(run_ptr(nb.send.start[ax]), (size_t) ((nb.send.end[ax] - nb.send.start[ax]) * plane), ncclFloat64, nb.rank, n.comm, st));
            else
                OPF_NCCL(n.Send(f->halo_send + soff[i], (size_t) (soff[i + 1] - soff[i]), ncclFloat64, nb.rank, n.comm, st));
        }
        for (int k = 0; k < nn; ++k) {
            const int i = rorder[k];
            const auto& nb = f->neighbors[i];
            if (direct[i]) OPF_NCCL(n.Recv(run_ptr(nb.recv.start[ax]), (size_t) ((nb.recv.end[ax] - nb.recv.start[ax]) * plane), ncclFloat64, nb.rank, n.comm, st));
            else if (roff[i + 1] - roff[i] > 0)
                OPF_NCCL(n.Recv(f->halo_recv + roff[i], (size_t) (roff[i + 1] - roff[i]), ncclFloat64, nb.rank, n.comm, st));
        }
        OPF_NCCL(n.GroupEnd());
        for (int i = 0; i < nn; ++i) {
            if (direct[i] || roff[i + 1] - roff[i] <= 0) continue;
            if (int rc = dense(f->neighbors[i].recv, f->halo_recv + roff[i], true)) return rc;
        }
        return OPF_OK;
    }
}// namespace opfe

extern "C" {

int opf_comm_unique_id(void* id128) {
    if (!id128) return opfe::fail(OPF_ERR_INVALID, "null id");
    if (int rc = load_nccl()) return rc;
    ncclUniqueId id;
    OPF_NCCL(nc().GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return OPF_OK;
}

int opf_comm_init(int rank, int n_ranks, const void* id128) {
    if (!id128 || rank < 0 || rank >= n_ranks) return opfe::fail(OPF_ERR_INVALID, "bad arguments");
    if (int rc = opfe::require_device()) return rc;
    if (int rc = load_nccl()) return rc;
    Nccl& n = nc();
    if (n.comm) return OPF_OK;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    OPF_NCCL(n.CommInitRank(&n.comm, n_ranks, id, rank));
    n.rank = rank;
    n.size = n_ranks;
    OPF_CUDA(cudaMalloc(&n.dbuf, sizeof(double) * 64));
    return OPF_OK;
}
int opf_comm_rank(void) { return nc().rank; }
int opf_comm_size(void) { return nc().size; }

int opf_comm_allreduce(double* values, int cnt, int rop) {
    Nccl& n = nc();
    if (cnt <= 0 || cnt > 64 || !values) return opfe::fail(OPF_ERR_INVALID, "allreduce count must be 1..64");
    if (!n.comm) return OPF_OK;// single process: identity
    opfe::Context& c = opfe::ctx();
    const int op = rop == OPF_RED_MAX || rop == OPF_RED_ABSMAX ? ncclMax : (rop == OPF_RED_MIN ? ncclMin : ncclSum);
    OPF_CUDA(cudaMemcpyAsync(n.dbuf, values, sizeof(double) * cnt, cudaMemcpyHostToDevice, c.stream));
    OPF_NCCL(n.AllReduce(n.dbuf, n.dbuf, (size_t) cnt, ncclFloat64, op, n.comm, c.stream));
    OPF_CUDA(cudaMemcpyAsync(values, n.dbuf, sizeof(double) * cnt, cudaMemcpyDeviceToHost, c.stream));
    OPF_CUDA(cudaStreamSynchronize(c.stream));
    return OPF_OK;
}

// CartesianField::resplitWithStrategy (CartesianField.hpp:83-177): the field keeps its values and moves to another decomposition.
// Every rank sends  old localRange ∩ new block of r  to r and receives  new localRange ∩ old block of r  (dense boxes through one
// staging buffer each way, one NCCL group), then takes over the storage planned for the new map and refreshes its padding.
// the field the move ends in: same description as f, blocks from split_map (a device-free plan, or a field with storage)
static opf_field_s* resplit_target(opf_field_s* f, const opf_range* split_map, bool plan_only) {
    using namespace opfe;
    opf_field_desc d{};
    d.mesh = f->mesh;
    for (int a = 0; a < f->dim; ++a) {
        d.loc[a] = f->loc[a];
        for (int sd = 0; sd < 2; ++sd) {
            if (f->bc[a][sd].face_dev || !f->bc[a][sd].face.empty()) {
                fail(OPF_ERR_UNSUPPORTED, "opf_field_resplit: field '%s' has functor boundary values (their face slabs belong to the old blocks)", f->name.c_str());
                return nullptr;
            }
            d.bc[a][sd].type = f->bc[a][sd].type;
            d.bc[a][sd].value = f->bc[a][sd].value;
            d.bc[a][sd].face = nullptr;
            d.ext[a][sd] = f->ext[a][sd];
        }
    }
    d.padding = f->padding;
    d.n_ranks = f->n_ranks;
    d.rank = f->rank;
    d.split_map = split_map;
    return plan_only ? opf_field_plan(&d, f->name.c_str()) : opf_field_create(&d, f->name.c_str());
}
// what rank `me` sends to / receives from every rank r:  old localRange ∩ new block of r  /  new localRange ∩ old block of r
static void resplit_boxes(const opf_field_s* f, const opf_field_s* g, std::vector<opfe::Range>& sbox, std::vector<opfe::Range>& rbox) {
    const int R = f->n_ranks;
    sbox.assign(R, opfe::Range());
    rbox.assign(R, opfe::Range());
    for (int r = 0; r < R; ++r) {
        sbox[r] = opfe::common(f->local, g->split_map[r]);
        rbox[r] = opfe::common(g->local, f->split_map[r]);
    }
}

int opf_field_resplit_plan(opf_field_t f, const opf_range* split_map, opf_range* send, opf_range* recv, opf_range* new_local) {
    using namespace opfe;
    if (!f || !split_map || !send || !recv) return fail(OPF_ERR_INVALID, "opf_field_resplit_plan: null argument");
    if (f->n_ranks <= 1) return fail(OPF_ERR_INVALID, "opf_field_resplit_plan: field '%s' is not decomposed", f->name.c_str());
    opf_field_s* g = resplit_target(f, split_map, true);
    if (!g) return OPF_ERR_UNSUPPORTED;
    std::vector<Range> sbox, rbox;
    resplit_boxes(f, g, sbox, rbox);
    for (int r = 0; r < f->n_ranks; ++r) send[r] = to_c(sbox[r]), recv[r] = to_c(rbox[r]);
    if (new_local) *new_local = to_c(g->local);
    opf_field_destroy(g);
    return OPF_OK;
}

int opf_field_resplit(opf_field_t f, const opf_range* split_map) {
    using namespace opfe;
    if (!f || !split_map) return fail(OPF_ERR_INVALID, "opf_field_resplit: null argument");
    if (!f->buf[0]) return fail(OPF_ERR_INVALID, "field '%s' is a plan (opf_field_plan): it has no device storage", f->name.c_str());
    Nccl& n = nc();
    if (f->n_ranks <= 1) return OPF_OK;// the reference's method only acts under MPI (:88)
    if (!n.comm) return fail(OPF_ERR_COMM, "field '%s' is decomposed over %d ranks but opf_comm_init was not called", f->name.c_str(), f->n_ranks);
    opf_field_s* g = resplit_target(f, split_map, false);// collective: its first updatePadding exchanges (zero) halos
    if (!g) return OPF_ERR_UNSUPPORTED;
    const int R = f->n_ranks, me = f->rank;
    cudaStream_t st = ctx().stream;
    std::vector<Range> sbox, rbox;
    resplit_boxes(f, g, sbox, rbox);
    std::vector<long long> soff(R + 1, 0), roff(R + 1, 0);
    for (int r = 0; r < R; ++r) {
        soff[r + 1] = soff[r] + std::max<long long>(0, sbox[r].count());
        roff[r + 1] = roff[r] + std::max<long long>(0, rbox[r].count());
    }
    double *ss = nullptr, *rs = nullptr;
    auto cleanup = [&](int rc) {
        if (ss) cudaFree(ss);
        if (rs) cudaFree(rs);
        if (rc) opf_field_destroy(g);
        return rc;
    };
    if (cudaMalloc(&ss, sizeof(double) * std::max<long long>(1, soff[R])) != cudaSuccess || cudaMalloc(&rs, sizeof(double) * std::max<long long>(1, roff[R])) != cudaSuccess)
        return cleanup(fail(OPF_ERR_CUDA, "opf_field_resplit: staging allocation failed"));
    auto dense = [&](opf_field_s* fld, const Range& r, double* stage, bool unpack) {
        const long long n0 = r.end[0] - r.start[0], n1 = r.end[1] - r.start[1];
        return dense_convert(fld, fld->cur, stage, r, n0, n0 * n1, unpack, st);
    };
    for (int r = 0; r < R; ++r)
        if (sbox[r].count() > 0)
            if (int rc = dense(f, sbox[r], ss + soff[r], false)) return cleanup(rc);
    if (sbox[me].count() != rbox[me].count()) return cleanup(fail(OPF_ERR_INVALID, "opf_field_resplit: inconsistent self block"));
    if (sbox[me].count() > 0)
        if (cudaMemcpyAsync(rs + roff[me], ss + soff[me], sizeof(double) * sbox[me].count(), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
            return cleanup(fail(OPF_ERR_CUDA, "opf_field_resplit: local copy failed"));
    if (n.GroupStart() != 0) return cleanup(fail(OPF_ERR_COMM, "ncclGroupStart failed"));
    for (int r = 0; r < R; ++r) {
        if (r == me) continue;
        if (sbox[r].count() > 0 && n.Send(ss + soff[r], (size_t) sbox[r].count(), ncclFloat64, r, n.comm, st) != 0) return cleanup(fail(OPF_ERR_COMM, "ncclSend failed"));
        if (rbox[r].count() > 0 && n.Recv(rs + roff[r], (size_t) rbox[r].count(), ncclFloat64, r, n.comm, st) != 0) return cleanup(fail(OPF_ERR_COMM, "ncclRecv failed"));
    }
    if (n.GroupEnd() != 0) return cleanup(fail(OPF_ERR_COMM, "ncclGroupEnd failed"));
    for (int r = 0; r < R; ++r)
        if (rbox[r].count() > 0)
            if (int rc = dense(g, rbox[r], rs + roff[r], true)) return cleanup(rc);
    if (cudaStreamSynchronize(st) != cudaSuccess) return cleanup(fail(OPF_ERR_CUDA, "opf_field_resplit: exchange failed: %s", cudaGetErrorString(cudaGetLastError())));
    cleanup(OPF_OK);
    repeat_cache_clear();
    field_adopt(f, g);
    opf_field_destroy(g);
    return field_update_padding(f);
}

int opf_comm_finalize(void) {
    Nccl& n = nc();
    if (n.comm) {
        cudaStreamSynchronize(opfe::ctx().stream);
        n.CommDestroy(n.comm);
        n.comm = nullptr;
        cudaFree(n.dbuf);
        n.dbuf = nullptr;
    }
    n.rank = 0;
    n.size = 1;
    return OPF_OK;
}

}// extern "C"
