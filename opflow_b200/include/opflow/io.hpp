// opflow/io.hpp -- stream writers of the examples (uf << Utils::TimeStamp(t) << u) over the field's host mirror.
// Reference: src/Utils/Writers/{FieldStream.hpp, TecplotASCIIStream.hpp, RawBinaryStream.hpp, HDF5Stream.hpp}.  I/O is outside
// the hot path: the Tecplot writer emits the reference's POINT-format zones; H5Stream has no HDF5 library to
// link against in this build, so it writes the same records as raw little-endian blocks (<name>.h5.raw + a text index).
#pragma once
#include "field.hpp"

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

    struct TecplotASCIIStream {
        std::string path;
        std::ofstream of;
        TimeStamp time;
        bool alwaysWriteMesh = true, writeMesh = true, separate = false;
        TecplotASCIIStream() = default;
        explicit TecplotASCIIStream(const std::string& p) : path(p), of(p) {}
        auto& operator<<(const TimeStamp& t) {
            time = t;
            return *this;
        }
        auto& alwaysWriteMeshes(bool o) {
            alwaysWriteMesh = o;
            return *this;
        }
        auto& fixedMesh() {
            alwaysWriteMesh = false;
            return *this;
        }
        auto& dumpToSeparateFile() {
            separate = true;
            return *this;
        }
        void close() { of.close(); }
        template <internal::FieldType F>
        auto& operator<<(const F& f) {
            constexpr int dim = F::dim;
            static const char* xn[3] = {"X", "Y", "Z"};
            of << "TITLE = \"Solution of " << f.name << "\"\nVARIABLES = ";
            for (int d = 0; d < dim; ++d) of << "\"" << xn[d] << "\", ";
            of << "\"" << f.name << "\"\nZONE T=\"t=" << time.time << "\" ";
            static const char* in[3] = {"I", "J", "K"};
            for (int d = 0; d < dim; ++d) of << in[d] << "=" << (f.localRange.end[d] - f.localRange.start[d]) << " ";
            of << "SOLUTIONTIME=" << time.time << " DATAPACKING=POINT\n";
            of << std::scientific;
            rangeFor_s(f.localRange, [&](auto&& i) {
                for (int d = 0; d < dim; ++d) of << (f.loc[d] == LocOnMesh::Corner ? f.mesh.x(d, i[d]) : f.mesh.x(d, i[d]) + .5 * f.mesh.dx(d, i[d])) << " ";
                of << f.evalAt(i) << "\n";
            });
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
