// opflow/io.hpp -- stream writers of the examples (uf << Utils::TimeStamp(t) << u) over the field's host mirror.
// Reference: src/Utils/Writers/{FieldStream.hpp, TecplotASCIIStream.hpp, RawBinaryStream.hpp, HDF5Stream.hpp}.  I/O is outside
// the hot path: the Tecplot writer emits the reference's BLOCK-format zones from the host mirror; the binary streams
// (RawBinaryOStream / RawBinaryIStream in the reference's .cart layout, H5Stream as a raw container -- no HDF5 library here) take
// asynchronous device snapshots and write from a background thread, so output cadence does not stall the time loop.
#pragma once
#include "field.hpp"
#include <condition_variable>
#include <cstring>
#include <deque>
#include <iomanip>
#include <memory>
#include <mutex>
#include <thread>

namespace OpFlow::Utils {
    struct TimeStamp {
        double time = 0;
        TimeStamp() = default;
        explicit TimeStamp(double t) : time(t) {}
        operator double() const { return time; }
    };

    namespace detail {
        template <typename F>
        void for_each_local(const F& f, const std::function<void(const typename F::IndexType&)>& fn) {
            rangeFor_s(f.localRange, fn);
        }
    }// namespace detail

    // TecplotASCIIStream (TecplotASCIIStream.hpp:36-215): same on-disk layout -- one ORDERED / BLOCK zone per `<<`, coordinates of
    // the first zone only when the mesh is fixed (default), values with 10 digits
    struct TecplotASCIIStream {
        std::string path;
        std::ofstream of;
        TimeStamp time;
        bool writeMesh = true, fixed_mesh = true, separate_file = false;
        TecplotASCIIStream() = default;
        explicit TecplotASCIIStream(const std::string& p) : path(p), of(p) {}
        auto& operator<<(const TimeStamp& t) {
            time = t;
            return *this;
        }
        // the reference's flag reads inverted (`fixed_mesh` true == write the coordinates in every zone, its default;
        // fixedMesh() clears it: TecplotASCIIStream.hpp:51,119) -- kept, so files have the same zones
        auto& fixedMesh() {
            fixed_mesh = false;
            return *this;
        }
        auto& dumpToSeparateFile() {
            separate_file = true;
            if (!fixed_mesh) fixed_mesh = true;
            return *this;
        }
        void close() { of.close(); }
        template <internal::FieldType F>
        auto& operator<<(const F& f) {
            constexpr int dim = F::dim;
            if (separate_file) {
                std::string filename = path, ext;
                if (auto dot = filename.rfind('.'); dot != std::string::npos) {
                    ext = filename.substr(dot);
                    filename.erase(dot);
                }
                of.close();
                of.open(filename + std::format("_{:.6f}", time.time) + ext, std::ofstream::out | std::ofstream::ate);
            }
            if (of.tellp() == 0) {
                of << std::format("TITLE = \"Solution of {} \"\n", f.name);
                static const char* vars[3] = {R"("X")", R"("X", "Y")", R"("X", "Y", "Z")"};
                of << std::format("VARIABLES = {}, \"{}\"\n", vars[dim - 1], f.name);
            }
            of << "ZONE\nZONETYPE = ORDERED DATAPACKING = BLOCK\n";
            static const char* in[3] = {"I", "J", "K"};
            for (int d = 0; d < dim; ++d) of << (d ? " " : "") << in[d] << " = " << (f.localRange.end[d] - f.localRange.start[d]);
            of << "\n" << std::scientific << std::setprecision(10);
            of << std::format("SOLUTIONTIME = {}\n", time.time);
            if (!writeMesh) {
                if constexpr (dim == 1) of << "VARSHARELIST=([1]=1)\n";
                else
                    of << std::format("VARSHARELIST=([1-{}]=1)\n", dim);
            } else {
                for (int k = 0; k < dim; ++k) {
                    if (f.loc[k] == LocOnMesh::Corner) rangeFor_s(f.localRange, [&](auto&& i) { of << f.mesh.x(k, i[k]) << "\n"; });
                    else
                        rangeFor_s(f.localRange, [&](auto&& i) { of << Math::mid(f.mesh.x(k, i[k]), f.mesh.x(k, i[k] + 1)) << "\n"; });
                }
            }
            rangeFor_s(f.localRange, [&](auto&& i) { of << f.evalAt(i) << "\n"; });
            of.flush();
            if (!separate_file) writeMesh = fixed_mesh;
            if (separate_file) close();
            return *this;
        }
    };

    // ---- binary streams with device-resident snapshots.  `stream << u` packs u's localRange on the device in stream
    // order (opf_field_snapshot: the values as of this statement), copies it to pinned host memory on a copy stream and hands the
    // file write to a background thread: the time loop continues at once.  close() / the destructor wait for the pending writes.
    inline constexpr unsigned StreamIn = 1u, StreamOut = 1u << 1, StreamASCII = 1u << 2, StreamBinary = 1u << 3;// StreamTrait.hpp flags
    namespace detail {
        struct AsyncWriter {// one worker per stream object, jobs run in submission order
            std::thread th;
            std::mutex mu;
            std::condition_variable cv;
            std::deque<std::function<void()>> q;
            bool stop = false, busy = false;
            void submit(std::function<void()> job) {
                std::unique_lock<std::mutex> lk(mu);
                if (!th.joinable()) th = std::thread([this] { loop(); });
                q.push_back(std::move(job));
                cv.notify_all();
            }
            void loop() {
                for (;;) {
                    std::function<void()> job;
                    {
                        std::unique_lock<std::mutex> lk(mu);
                        cv.wait(lk, [&] { return stop || !q.empty(); });
                        if (q.empty()) return;
                        job = std::move(q.front());
                        q.pop_front();
                        busy = true;
                    }
                    job();
                    {
                        std::unique_lock<std::mutex> lk(mu);
                        busy = false;
                        cv.notify_all();
                    }
                }
            }
            void drain() {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return q.empty() && !busy; });
            }
            ~AsyncWriter() {
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return q.empty() && !busy; });
                    stop = true;
                    cv.notify_all();
                }
                if (th.joinable()) th.join();
            }
        };
        template <typename T>
        void put(std::string& b, const T& v) {
            b.append(reinterpret_cast<const char*>(&v), sizeof v);
        }
        // async snapshot of f's localRange -> a job that writes `header` followed by the values to `file` (mode "wb" / "ab")
        template <typename F>
        void snapshot_to_file(AsyncWriter& w, const F& f, std::string file, const char* mode, std::string header) {
            f.syncToDevice();
            const long long n = f.localRange.count();
            double* host = static_cast<double*>(opf_host_alloc(sizeof(double) * (unsigned long long) std::max<long long>(n, 1)));
            OpFlow::internal::check_ptr(host, "opf_host_alloc");
            const opf_range r = OpFlow::internal::to_c(f.localRange);
            opf_snapshot_t snap = opf_field_snapshot(f.h, &r, host);
            OpFlow::internal::check_ptr(snap, "opf_field_snapshot");
            w.submit([=, header = std::move(header), file = std::move(file)] {
                opf_snapshot_wait(snap);
                if (FILE* fp = std::fopen(file.c_str(), mode)) {
                    std::fwrite(header.data(), 1, header.size(), fp);
                    std::fwrite(host, sizeof(double), (std::size_t) n, fp);
                    std::fclose(fp);
                } else
                    std::fprintf(stderr, "[opflow] cannot open %s for writing\n", file.c_str());
                opf_host_free(host);
            });
        }
        inline std::string parent_dir(const std::string& p) {
            const auto slash = p.rfind('/');
            return slash == std::string::npos ? std::string(".") : (slash == 0 ? std::string("/") : p.substr(0, slash));
        }
    }// namespace detail

    // RawBinaryOStream / RawBinaryIStream (RawBinaryStream.hpp:43-227): the reference's on-disk layout, one file per field and output
    // count -- <dir>/<name>_<count>[_<rank>].cart =
    //   int name_len | name | int dim | int nproc | double time | dim x (int, int) mesh range | mesh coordinates per axis |
    //   dim x (int, int) accessibleRange | dim x (int, int) localRange | localRange values, axis 0 fastest
    struct RawBinaryOStream {
        std::string dir;
        TimeStamp time;
        int count = 0;
        std::unique_ptr<detail::AsyncWriter> writer = std::make_unique<detail::AsyncWriter>();
        RawBinaryOStream() = default;
        explicit RawBinaryOStream(const std::string& path) : dir(detail::parent_dir(path)) {}
        auto& operator<<(const TimeStamp& t) {
            time = t;
            return *this;
        }
        void setCounterTo(int c) { count = c; }
        auto& fixedMesh() { return *this; }
        auto& dumpToSeparateFile() { return *this; }
        void close() { writer->drain(); }
        template <OpFlow::internal::FieldType F>
        auto& operator<<(const F& f) {
            constexpr int dim = F::dim;
            const int nproc = getWorkerCount(), rank = getWorkerId();
            const std::string file = nproc > 1 ? std::format("{}/{}_{}_{}.cart", dir, f.name, count, rank) : std::format("{}/{}_{}.cart", dir, f.name, count);
            std::string h;
            detail::put(h, (int) f.name.size());
            h += f.name;
            detail::put(h, dim);
            detail::put(h, nproc);
            detail::put(h, (double) time.time);
            const auto mr = f.mesh.getRange();
            for (int d = 0; d < dim; ++d) detail::put(h, mr.start[d]), detail::put(h, mr.end[d]);
            for (int d = 0; d < dim; ++d)
                for (int j = mr.start[d]; j < mr.end[d]; ++j) detail::put(h, (double) f.mesh.x(d, j));
            for (int d = 0; d < dim; ++d) detail::put(h, f.accessibleRange.start[d]), detail::put(h, f.accessibleRange.end[d]);
            for (int d = 0; d < dim; ++d) detail::put(h, f.localRange.start[d]), detail::put(h, f.localRange.end[d]);
            detail::snapshot_to_file(*writer, f, file, "wb", std::move(h));
            ++count;
            return *this;
        }
    };
    using RawBinaryStream = RawBinaryOStream;// the spelling of round 1's stub

    struct RawBinaryIStream {
        std::string dir;
        int count = 0;
        RawBinaryIStream() = default;
        explicit RawBinaryIStream(const std::string& path) : dir(detail::parent_dir(path)) {}
        void setCounterTo(int c) { count = c; }
        template <OpFlow::internal::FieldType F>
        auto& operator>>(F& f) {// checks the meta data like the reference, then uploads the block and refreshes the padding
            constexpr int dim = F::dim;
            const int nproc = getWorkerCount(), rank = getWorkerId();
            const std::string file = nproc > 1 ? std::format("{}/{}_{}_{}.cart", dir, f.name, count, rank) : std::format("{}/{}_{}.cart", dir, f.name, count);
            FILE* fp = std::fopen(file.c_str(), "rb");
            if (!fp) {
                OP_CRITICAL("RawBinaryIStream: cannot open {}", file);
                OP_ABORT;
            }
            auto get_int = [&] {
                int v = 0;
                if (std::fread(&v, sizeof v, 1, fp) != 1) v = -1;
                return v;
            };
            const int name_len = get_int();
            std::string name((std::size_t) std::max(0, name_len), ' ');
            if (name_len > 0 && std::fread(name.data(), 1, (std::size_t) name_len, fp) != (std::size_t) name_len) name.clear();
            if (name != f.name) OP_WARN("Field's name {} in file is different from dst field {}", name, f.name);
            bool ok = get_int() == dim && get_int() == nproc;
            double t = 0;
            ok = ok && std::fread(&t, sizeof t, 1, fp) == 1;
            const auto mr = f.mesh.getRange();
            for (int d = 0; d < dim && ok; ++d) ok = get_int() == mr.start[d] && get_int() == mr.end[d];
            for (int d = 0; d < dim && ok; ++d)
                for (int j = mr.start[d]; j < mr.end[d] && ok; ++j) {
                    double x = 0;
                    ok = std::fread(&x, sizeof x, 1, fp) == 1 && x == f.mesh.x(d, j);
                }
            for (int d = 0; d < dim && ok; ++d) ok = get_int() == f.accessibleRange.start[d] && get_int() == f.accessibleRange.end[d];
            for (int d = 0; d < dim && ok; ++d) ok = get_int() == f.localRange.start[d] && get_int() == f.localRange.end[d];
            if (!ok) {
                OP_CRITICAL("Field read error: meta data of {} does not match field {}", file, f.name);
                OP_ABORT;
            }
            std::vector<double> buf((std::size_t) f.localRange.count());
            if (std::fread(buf.data(), sizeof(double), buf.size(), fp) != buf.size()) {
                OP_CRITICAL("Field read error: {} is truncated", file);
                OP_ABORT;
            }
            std::fclose(fp);
            const opf_range r = OpFlow::internal::to_c(f.localRange);
            OpFlow::internal::check_rc(opf_field_upload(f.h, &r, buf.data()), "opf_field_upload");
            f.touch();
            f.updatePadding();
            ++count;
            return *this;
        }
    };

    // H5Stream (HDF5Stream.hpp:47-363).  There is no HDF5 library in this build -- where the reference itself, built without
    // OPFLOW_WITH_HDF5, turns the stream into a warning ("H5Stream not enabled", HDF5Stream.hpp:107,127) -- so the same records
    // (/T=<time>/<name>: local block + ranges) go to a self-describing raw container next to the requested name:
    //   <path>[.rank<r>]  =  sequence of  "OPF5" | double time | int name_len | name | int dim | dim x (int, int) accessibleRange |
    //                        dim x (int, int) localRange | localRange values, axis 0 fastest
    // one file per rank (ranks never share a file), written through asynchronous device snapshots; StreamIn reads it back for restarts.
    struct H5Stream {
        std::string path;
        unsigned mode = StreamOut;
        TimeStamp time;
        bool separate_file = false, truncated = false;
        std::unique_ptr<detail::AsyncWriter> writer = std::make_unique<detail::AsyncWriter>();
        long read_pos = 0;
        H5Stream() = default;
        H5Stream(const H5Stream&) = delete;
        H5Stream(H5Stream&&) noexcept = default;
        explicit H5Stream(const std::string& p, unsigned m = StreamOut) : path(p), mode(m) {
            static bool warned = false;
            if (!warned && getWorkerId() == 0) {
                OP_WARN("H5Stream: no HDF5 library in this build -- '{}' is written as an OpFlow-B200 raw container (same records, one file per rank)", p);
                warned = true;
            }
        }
        ~H5Stream() { close(); }
        auto& operator<<(const TimeStamp& t) {
            time = t;
            return *this;
        }
        auto& fixedMesh() { return *this; }
        auto& dumpToSeparateFile() {
            separate_file = true;
            return *this;
        }
        auto& moveToTime(const TimeStamp& t) {
            time = t;
            return *this;
        }
        void close() {
            if (writer) writer->drain();
        }
        std::string filename() const {
            std::string f = path;
            if (separate_file) {
                std::string ext;
                if (auto dot = f.rfind('.'); dot != std::string::npos && dot > f.rfind('/') + 1) {
                    ext = f.substr(dot);
                    f.erase(dot);
                }
                f += std::format("_{:.6f}", time.time) + ext;
            }
            if (getWorkerCount() > 1) f += std::format(".rank{}", getWorkerId());
            return f;
        }
        template <OpFlow::internal::FieldType F>
        auto& operator<<(const F& f) {
            constexpr int dim = F::dim;
            std::string h = "OPF5";
            detail::put(h, (double) time.time);
            detail::put(h, (int) f.name.size());
            h += f.name;
            detail::put(h, dim);
            for (int d = 0; d < dim; ++d) detail::put(h, f.accessibleRange.start[d]), detail::put(h, f.accessibleRange.end[d]);
            for (int d = 0; d < dim; ++d) detail::put(h, f.localRange.start[d]), detail::put(h, f.localRange.end[d]);
            const bool fresh = separate_file || !truncated;// first record of a run replaces an old file, later ones append
            truncated = true;
            detail::snapshot_to_file(*writer, f, filename(), fresh ? "wb" : "ab", std::move(h));
            return *this;
        }
        // reads the record of `f.name` at the stream's current time (moveToTime / << TimeStamp), or the next record of that name
        template <OpFlow::internal::FieldType F>
        auto& operator>>(F& f) {
            constexpr int dim = F::dim;
            FILE* fp = std::fopen(filename().c_str(), "rb");
            if (!fp) {
                OP_CRITICAL("H5Stream: cannot open {}", filename());
                OP_ABORT;
            }
            bool found = false;
            std::vector<double> buf;
            for (;;) {
                char magic[4];
                double t = 0;
                int name_len = 0, fdim = 0;
                if (std::fread(magic, 1, 4, fp) != 4 || std::memcmp(magic, "OPF5", 4) != 0) break;
                if (std::fread(&t, sizeof t, 1, fp) != 1 || std::fread(&name_len, sizeof name_len, 1, fp) != 1) break;
                std::string name((std::size_t) name_len, ' ');
                if (std::fread(name.data(), 1, name.size(), fp) != name.size() || std::fread(&fdim, sizeof fdim, 1, fp) != 1) break;
                std::vector<int> rg((std::size_t) 4 * fdim);
                if (std::fread(rg.data(), sizeof(int), rg.size(), fp) != rg.size()) break;
                long long n = 1;
                for (int d = 0; d < fdim; ++d) n *= rg[2 * fdim + 2 * d + 1] - rg[2 * fdim + 2 * d];
                bool match = name == f.name && fdim == dim && t == time.time;
                for (int d = 0; d < dim && match; ++d) match = rg[2 * dim + 2 * d] == f.localRange.start[d] && rg[2 * dim + 2 * d + 1] == f.localRange.end[d];
                if (match) {
                    buf.resize((std::size_t) n);
                    found = std::fread(buf.data(), sizeof(double), buf.size(), fp) == buf.size();
                    break;
                }
                std::fseek(fp, (long) (n * (long long) sizeof(double)), SEEK_CUR);
            }
            std::fclose(fp);
            if (!found) {
                OP_CRITICAL("H5Stream: no record /T={}/{} with the field's local range in {}", time.time, f.name, filename());
                OP_ABORT;
            }
            const opf_range r = OpFlow::internal::to_c(f.localRange);
            OpFlow::internal::check_rc(opf_field_upload(f.h, &r, buf.data()), "opf_field_upload");
            f.touch();
            f.updatePadding();
            return *this;
        }
    };
}// namespace OpFlow::Utils
