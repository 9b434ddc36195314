// opflow/parallel.hpp -- environment, parallel plan and split strategies of the B200 front-end.
// Reference: src/Core/Environment.hpp:29-90, src/Core/Parallel/ParallelInfo.hpp, ParallelPlan.hpp:21-50,
// AbstractSplitStrategy.hpp:24-32, EvenSplitStrategy.hpp:57-192, ManualSplitStrategy.hpp:34-57.
// "Distributed workers" are processes, one per GPU, launched by torchrun / mpirun / srun (RANK, WORLD_SIZE, LOCAL_RANK or the
// OMPI_/PMI_ equivalents).  MPI_Init becomes: pick the GPU, create the NCCL communicator (id exchanged through a file).
#pragma once
#include "field.hpp"
#include <cctype>
#include <cstring>
#include <ctime>
#include <fcntl.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>

namespace OpFlow {
    using ParallelType = unsigned;
    namespace ParallelIdentifier {
        inline constexpr ParallelType None = 0, DistributeMem = 0x1, SharedMem = 0x2, Heterogeneous = 0x4;
    }
    enum class DistributeMemType { None, MPI };
    enum class SharedMemType { None, OpenMP, TBB };
    enum class HeterogeneousType { None, CUDA };
    struct NodeInfo {
        DistributeMemType type = DistributeMemType::None;
        int node_count = 1;
    };
    struct ThreadInfo {
        SharedMemType type = SharedMemType::None;
        int thread_count = 1;
    };
    struct DeviceInfo {
        HeterogeneousType type = HeterogeneousType::CUDA;
        int device_count = 1;
    };
    struct ParallelInfo {
        ParallelType parallelType = 0;
        NodeInfo nodeInfo {};
        ThreadInfo threadInfo {};
        DeviceInfo deviceInfo {};
    };
    struct ParallelPlan {
        ParallelInfo info;
        int distributed_workers_count = 1;
        int shared_memory_workers_count = 1;
        int heterogeneous_workers_count = 0;
        [[nodiscard]] bool serialMode() const { return distributed_workers_count == 1 && shared_memory_workers_count == 1 && heterogeneous_workers_count == 0; }
        [[nodiscard]] bool singleNodeMode() const { return distributed_workers_count == 1; }
        [[nodiscard]] bool multiThreadMode() const { return distributed_workers_count == 1 && shared_memory_workers_count > 1 && heterogeneous_workers_count == 0; }
        [[nodiscard]] bool deviceMode() const { return heterogeneous_workers_count > 0; }
    };

    namespace internal {
        inline ParallelInfo GLOBAL_PARALLELINFO;
        inline ParallelPlan GLOBAL_PARALLELPLAN;
        inline int env_int(std::initializer_list<const char*> names, int dflt) {
            for (const char* n : names)
                if (const char* v = std::getenv(n)) return std::atoi(v);
            return dflt;
        }
        inline int env_rank() { return env_int({"RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", "SLURM_PROCID"}, 0); }
        inline int env_world() { return env_int({"WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", "SLURM_NTASKS"}, 1); }
        inline int env_local_rank() { return env_int({"LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", "SLURM_LOCALID"}, env_rank()); }
        // what every rank of ONE job shares and two jobs do not: an explicit id, the elastic run id, the batch job id, or -- for the
        // ranks torchrun / mpirun fork on one node -- the launcher's process id
        inline std::string job_nonce() {
            for (const char* n : {"OPF_JOB_ID", "TORCHELASTIC_RUN_ID", "SLURM_JOB_ID", "PMIX_NAMESPACE"})
                if (const char* v = std::getenv(n))
                    if (*v && std::string(v) != "none") {
                        std::string out;
                        for (const char* c = v; *c && out.size() < 40; ++c) out.push_back(std::isalnum((unsigned char) *c) ? *c : '-');
                        return out;
                    }
            return "ppid" + std::to_string((long) ::getppid());
        }
        struct RendezvousRecord {
            char magic[8];
            char nonce[48];
            long long stamp;
            unsigned char id[128];
        };
    }// namespace internal

    inline int getWorkerId() { return opf_comm_rank(); }
    inline int getWorkerCount() { return opf_comm_size(); }

    // InitEnvironment (Environment.hpp:29-33): MPI_Init -> GPU selection + NCCL communicator
    inline void InitEnvironment(int*, char***) {
        const int world = internal::env_world(), rank = internal::env_rank();
        internal::check_rc(opf_init(world > 1 ? internal::env_local_rank() : -1), "opf_init");
        if (world <= 1) return;
        // rendezvous: rank 0 publishes the ncclUniqueId in a file every rank of the job can see.  The record carries a per-job
        // nonce and rank 0's publication time, the file is created exclusively (O_EXCL | O_NOFOLLOW, mode 0600) under a name that
        // includes the user id and the nonce, and readers only accept a regular file owned by themselves whose nonce matches and
        // whose time stamp is not older than their own start: a stale file left by a crashed job is neither read nor followed.
        // Multi-node jobs (no shared /tmp) must point OPF_RENDEZVOUS_DIR at a shared directory and set OPF_JOB_ID.
        const char* dir = std::getenv("OPF_RENDEZVOUS_DIR");
        const std::string nonce = internal::job_nonce();
        const std::string path = std::string(dir ? dir : "/tmp") + "/opflow_b200_nccl_u" + std::to_string((long) ::geteuid()) + "_" + nonce + "_"
                                 + std::to_string(internal::env_int({"MASTER_PORT"}, 0)) + "_" + std::to_string(world) + ".id";
        internal::RendezvousRecord rec{};
        unsigned char id[128];
        const long long t_start = (long long) ::time(nullptr);
        if (rank == 0) {
            internal::check_rc(opf_comm_unique_id(id), "opf_comm_unique_id");
            std::memcpy(rec.magic, "OPFNCCL1", 8);
            std::snprintf(rec.nonce, sizeof rec.nonce, "%s", nonce.c_str());
            rec.stamp = (long long) ::time(nullptr);
            std::memcpy(rec.id, id, sizeof id);
            const std::string tmp = path + ".tmp";
            ::unlink(path.c_str());// a stale record of a crashed job with the same name
            ::unlink(tmp.c_str());
            const int fd = ::open(tmp.c_str(), O_WRONLY | O_CREAT | O_EXCL | O_NOFOLLOW, 0600);
            if (fd < 0 || ::write(fd, &rec, sizeof rec) != (ssize_t) sizeof rec) {
                OP_CRITICAL("InitEnvironment: cannot publish the NCCL id at {}", tmp);
                OP_ABORT;
            }
            ::close(fd);
            if (std::rename(tmp.c_str(), path.c_str()) != 0) {
                OP_CRITICAL("InitEnvironment: cannot rename {} -> {}", tmp, path);
                OP_ABORT;
            }
        } else {
            for (int tries = 0;; ++tries) {
                bool ok = false;
                const int fd = ::open(path.c_str(), O_RDONLY | O_NOFOLLOW);
                if (fd >= 0) {
                    struct stat st {};
                    if (::fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_uid == ::geteuid() && st.st_size == (off_t) sizeof rec
                        && ::read(fd, &rec, sizeof rec) == (ssize_t) sizeof rec && std::memcmp(rec.magic, "OPFNCCL1", 8) == 0
                        && nonce == std::string(rec.nonce, ::strnlen(rec.nonce, sizeof rec.nonce)) && rec.stamp >= t_start - 600)
                        ok = true;
                    ::close(fd);
                }
                if (ok) break;
                if (tries > 12000) {
                    OP_CRITICAL("InitEnvironment: no valid NCCL id at {} after 120 s (multi-node: set OPF_RENDEZVOUS_DIR to a shared directory and OPF_JOB_ID)", path);
                    OP_ABORT;
                }
                std::this_thread::sleep_for(std::chrono::milliseconds(10));
            }
            std::memcpy(id, rec.id, sizeof id);
        }
        internal::check_rc(opf_comm_init(rank, world, id), "opf_comm_init");
        double one = 1.0;// doubles as a barrier: every rank has joined before rank 0 removes the file
        internal::check_rc(opf_comm_allreduce(&one, 1, OPF_RED_SUM), "opf_comm_allreduce");
        if (rank == 0) std::remove(path.c_str());
    }
    inline void FinalizeEnvironment() {
        opf_synchronize();
        opf_comm_finalize();
    }
    struct EnvironmentGardian {
        EnvironmentGardian(int* argc, char*** argv) { InitEnvironment(argc, argv); }
        ~EnvironmentGardian() { FinalizeEnvironment(); }
    };

    inline auto makeParallelInfo() {
        ParallelInfo ret;
        const int world = opf_comm_size();
        if (world > 1) {
            ret.parallelType |= ParallelIdentifier::DistributeMem;
            ret.nodeInfo.type = DistributeMemType::MPI;
        }
        ret.nodeInfo.node_count = world;
        ret.parallelType |= ParallelIdentifier::SharedMem | ParallelIdentifier::Heterogeneous;
        ret.threadInfo.type = SharedMemType::None;
        ret.threadInfo.thread_count = 1;// host threads play no role: the sweep runs on the GPU
        ret.deviceInfo.type = HeterogeneousType::CUDA;
        ret.deviceInfo.device_count = 1;
        return ret;
    }
    inline ParallelPlan makeParallelPlan(ParallelInfo info, ParallelType pbit) {// ParallelPlan.hpp:38-50
        const bool dist_bit = pbit & ParallelIdentifier::DistributeMem, sm_bit = pbit & ParallelIdentifier::SharedMem,
                   device_bit = pbit & ParallelIdentifier::Heterogeneous;
        ParallelPlan ret;
        ret.info = info;
        ret.distributed_workers_count = dist_bit ? info.nodeInfo.node_count : 1;
        ret.shared_memory_workers_count = sm_bit ? info.threadInfo.thread_count : 1;
        ret.heterogeneous_workers_count = device_bit ? info.deviceInfo.device_count : 0;
        return ret;
    }
    inline auto& getGlobalParallelInfo() { return internal::GLOBAL_PARALLELINFO; }
    inline ParallelPlan& getGlobalParallelPlan() { return internal::GLOBAL_PARALLELPLAN; }
    inline void setGlobalParallelInfo(const ParallelInfo& info) { internal::GLOBAL_PARALLELINFO = info; }
    inline void setGlobalParallelPlan(const ParallelPlan& plan) {
        internal::GLOBAL_PARALLELPLAN = plan;
        internal::g_host_threads = plan.shared_memory_workers_count > 0 ? plan.shared_memory_workers_count : 1;// host rangeFor workers
    }

    // EvenSplitStrategy<F>::getSplitMap (EvenSplitStrategy.hpp:57-192): the arithmetic lives in the engine (opf_split_even),
    // bit-exact against EvenSplitStrategyTest.cpp:28-134
    template <typename F>
    struct EvenSplitStrategy : AbstractSplitStrategy<F> {
        using Range = typename F::RangeType;
        std::string strategyName() const override { return "EvenSplit"; }
        std::vector<Range> getSplitMap(const Range& range, const ParallelPlan& plan) override {
            const int n = plan.distributed_workers_count;
            std::vector<opf_range> out(n);
            const opf_range r = internal::to_c(range);
            internal::check_rc(opf_split_even(F::dim, &r, n, out.data()), "opf_split_even");
            std::vector<Range> ret;
            for (auto& o : out) ret.push_back(internal::from_c<F::dim>(o));
            return ret;
        }
        Range splitRange(const Range& range, const ParallelPlan& plan) override { return getSplitMap(range, plan)[getWorkerId()]; }
    };
    // slabs along the slowest axis: one contiguous face per neighbour, exchange overlapped with the interior sweep
    template <typename F>
    struct SlabSplitStrategy : AbstractSplitStrategy<F> {
        using Range = typename F::RangeType;
        std::string strategyName() const override { return "SlabSplit"; }
        std::vector<Range> getSplitMap(const Range& range, const ParallelPlan& plan) override {
            const int n = plan.distributed_workers_count;
            std::vector<opf_range> out(n);
            const opf_range r = internal::to_c(range);
            internal::check_rc(opf_split_slab(F::dim, &r, n, out.data()), "opf_split_slab");
            std::vector<Range> ret;
            for (auto& o : out) ret.push_back(internal::from_c<F::dim>(o));
            return ret;
        }
        Range splitRange(const Range& range, const ParallelPlan& plan) override { return getSplitMap(range, plan)[getWorkerId()]; }
    };
    // ManualSplitStrategy (ManualSplitStrategy.hpp:34-57): the caller supplies the cell-centred block of every rank
    template <typename F>
    struct ManualSplitStrategy : AbstractSplitStrategy<F> {
        using Range = typename F::RangeType;
        std::vector<Range> splitMap;
        std::string strategyName() const override { return "ManualSplit"; }
        std::vector<Range> getSplitMap(const Range&, const ParallelPlan&) override { return splitMap; }
        Range splitRange(const Range&, const ParallelPlan&) override { return splitMap[getWorkerId()]; }
    };

    // globalReduce (RangeFor.hpp:125-135): local reduce, gather one value per rank, fold in rank order
    template <std::size_t d, typename ReOp, typename Fn>
    auto globalReduce(const DS::Range<d>& range, ReOp&& op, Fn&& func) {
        auto local = rangeReduce_s(range, op, func);
        const int world = opf_comm_size();
        if (world <= 1) return local;
        using R = decltype(local);
        if (world > 64) {
            OP_CRITICAL("globalReduce supports up to 64 ranks");
            OP_ABORT;
        }
        double slots[64] = {0};// allgather as a sum of one-hot vectors: exact, every slot has one non-zero contribution
        slots[opf_comm_rank()] = static_cast<double>(local);
        internal::check_rc(opf_comm_allreduce(slots, world, OPF_RED_SUM), "opf_comm_allreduce");
        R acc = static_cast<R>(slots[0]);
        for (int r = 1; r < world; ++r) acc = op(acc, static_cast<R>(slots[r]));
        return acc;
    }
}// namespace OpFlow
