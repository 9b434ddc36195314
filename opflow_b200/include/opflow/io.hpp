// opflow/io.hpp -- stream writers of the examples (uf << Utils::TimeStamp(t) << u) over the field's host mirror.
// Reference: src/Utils/Writers/{FieldStream.hpp, TecplotASCIIStream.hpp, RawBinaryStream.hpp, HDF5Stream.hpp}.  I/O is outside
// the hot path: the Tecplot writer emits the reference's POINT-format zones; H5Stream has no HDF5 library to
// link against in this build, so it writes the same records as raw little-endian blocks (<name>.h5.raw + a text index).
#pragma once
#include "field.hpp"
#include <iomanip>

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

    struct RawBinaryStream {
        std::string path;
        std::ofstream of;
        TimeStamp time;
        RawBinaryStream() = default;
        explicit RawBinaryStream(const std::string& p) : path(p), of(p, std::ios::binary) {}
        auto& operator<<(const TimeStamp& t) {
            time = t;
            return *this;
        }
        auto& fixedMesh() { return *this; }
        auto& dumpToSeparateFile() { return *this; }
        void close() { of.close(); }
        template <internal::FieldType F>
        auto& operator<<(const F& f) {// x-fastest local block, like RawBinaryOStream (RawBinaryStream.hpp:98-227)
            rangeFor_s(f.localRange, [&](auto&& i) {
                const double v = f.evalAt(i);
                of.write(reinterpret_cast<const char*>(&v), sizeof v);
            });
            return *this;
        }
    };

    struct H5Stream {
        std::string path;
        std::ofstream of, index;
        TimeStamp time;
        H5Stream() = default;
        explicit H5Stream(const std::string& p, unsigned = 0) : path(p), of(p + ".raw", std::ios::binary), index(p + ".index") {}
        auto& operator<<(const TimeStamp& t) {
            time = t;
            return *this;
        }
        auto& fixedMesh() { return *this; }
        auto& dumpToSeparateFile() { return *this; }
        void close() {
            of.close();
            index.close();
        }
        template <internal::FieldType F>
        auto& operator<<(const F& f) {
            index << "/T=" << time.time << "/" << f.name << " offset=" << of.tellp() << " extents=";
            for (int d = 0; d < F::dim; ++d) index << (f.localRange.end[d] - f.localRange.start[d]) << (d + 1 < F::dim ? "x" : "\n");
            rangeFor_s(f.localRange, [&](auto&& i) {
                const double v = f.evalAt(i);
                of.write(reinterpret_cast<const char*>(&v), sizeof v);
            });
            return *this;
        }
    };
}// namespace OpFlow::Utils
