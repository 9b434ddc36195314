// oracle/ref_drivers/ref_common.hpp -- TEST INFRASTRUCTURE (drives the unmodified reference).
// Our own helper code; it only *calls* the reference's public API (OpFlow headers).
#pragma once
#include <OpFlow>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace refdrv {
    using namespace OpFlow;

    // OPFD dump: "OPFD" | int32 dim | int32 start[dim] | int32 end[dim] | float64 data (axis 0 fastest)
    template <typename F>
    void dump_field(FILE* fp, const F& f, bool with_ghosts) {
        constexpr int dim = OpFlow::internal::FieldExprTrait<F>::dim;
        auto r = with_ghosts ? f.getLocalReadableRange() : f.localRange;
        int32_t d = dim;
        fwrite("OPFD", 1, 4, fp);
        fwrite(&d, 4, 1, fp);
        for (int i = 0; i < dim; ++i) { int32_t s = r.start[i]; fwrite(&s, 4, 1, fp); }
        for (int i = 0; i < dim; ++i) { int32_t e = r.end[i]; fwrite(&e, 4, 1, fp); }
        std::vector<double> buf;
        buf.reserve(r.count());
        rangeFor_s(r, [&](auto&& i) { buf.push_back(f.evalAt(i)); });
        fwrite(buf.data(), 8, buf.size(), fp);
    }

    template <typename F>
    void dump_field(const std::string& path, const F& f, bool with_ghosts) {
        FILE* fp = fopen(path.c_str(), "wb");
        if (!fp) { perror(path.c_str()); exit(2); }
        dump_field(fp, f, with_ghosts);
        fclose(fp);
    }

    inline void set_threads(int nt) {
        auto info = makeParallelInfo();
        info.threadInfo.thread_count = nt;
        setGlobalParallelInfo(info);
        setGlobalParallelPlan(makeParallelPlan(getGlobalParallelInfo(), ParallelIdentifier::SharedMem));
    }

    inline double now() {
        return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    }

    inline const char* arg(int argc, char** argv, const char* key, const char* def) {
        for (int i = 1; i + 1 < argc; ++i)
            if (!strcmp(argv[i], key)) return argv[i + 1];
        return def;
    }
}// namespace refdrv
