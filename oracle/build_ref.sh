#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE, not product code.
#
# Builds the *unmodified reference* (OpFlow, header-only C++23/26) plus the two vendored
# third-party libraries its hot path needs (oneTBB 2022.1 for rangeFor, HYPRE 2.33.0 for the
# implicit solve) into oracle/_ref/ so that the driver programs under oracle/ref_drivers/
# (our own code, written against the reference's public API) can produce reference outputs.
#
#   * sources are read where they lie under $OPF_REFERENCE (default /root/reference);
#     nothing from the reference is copied into the repository -- the patched header copy
#     lives in a scratch dir under /tmp, only binaries/.so land in oracle/_ref/ (git-ignored).
#   * the reference's own build system (its top-level CMakeLists) is NOT run; cmake is only
#     used on the vendored third-party TBB/HYPRE trees, exactly as SURVEY.md section 8c verified.
#   * two-line patch: src/Core/Meta.hpp:55,64 use C++26 pack indexing (T...[0]) which g++ 13
#     lacks -> std::tuple_element_t.  Shims: <print> (std::print over std::format), Version.hpp.
#
# Usage: oracle/build_ref.sh [driver ...]     (no args = all drivers)
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${OPF_REFERENCE:-/root/reference}"
SCRATCH="${OPF_REF_SCRATCH:-/tmp/opflow_ref_build}"
OUT="$HERE/_ref"
JOBS="${OPF_JOBS:-$(nproc)}"
export CC=/usr/bin/gcc CXX=/usr/bin/g++   # $CC may point at a wrapper without libgomp

if [ ! -d "$REF/src/Core" ]; then
    echo "build_ref: reference tree not found at $REF (fine on the GPU box: prebuilt oracle/_ref travels)"
    exit 0
fi
mkdir -p "$SCRATCH" "$OUT/bin" "$OUT/lib"

# 1. patched header copy (scratch only)
if [ ! -f "$SCRATCH/patched/.done" ]; then
    rm -rf "$SCRATCH/patched"; mkdir -p "$SCRATCH/patched"
    cp -r "$REF/src" "$REF/include" "$SCRATCH/patched/"
    sed -i 's|using type = T\.\.\.\[0\];|using type = std::tuple_element_t<0, std::tuple<T...>>;|; s|using type = T\.\.\.\[sizeof\.\.\.(T) - 1\];|using type = std::tuple_element_t<sizeof...(T) - 1, std::tuple<T...>>;|' \
        "$SCRATCH/patched/src/Core/Meta.hpp"
    grep -q '#include <tuple>' "$SCRATCH/patched/src/Core/Meta.hpp" || sed -i '0,/#include </s//#include <tuple>\n#include </' "$SCRATCH/patched/src/Core/Meta.hpp"
    touch "$SCRATCH/patched/.done"
fi

# 2. shims
mkdir -p "$SCRATCH/shim"
cat > "$SCRATCH/shim/print" <<'EOF'
#pragma once
#include <cstdio>
#include <format>
#include <string>
namespace std {
    template <typename... A> void print(std::format_string<A...> f, A&&... a) { auto s = std::format(f, std::forward<A>(a)...); std::fwrite(s.data(), 1, s.size(), stdout); }
    template <typename... A> void print(FILE* fp, std::format_string<A...> f, A&&... a) { auto s = std::format(f, std::forward<A>(a)...); std::fwrite(s.data(), 1, s.size(), fp); }
    template <typename... A> void println(std::format_string<A...> f, A&&... a) { auto s = std::format(f, std::forward<A>(a)...); s.push_back('\n'); std::fwrite(s.data(), 1, s.size(), stdout); }
}
EOF
cat > "$SCRATCH/shim/Version.hpp" <<'EOF'
#pragma once
#include <string_view>
namespace OpFlow::internal { inline constexpr auto OPFLOW_VERSION_STRING = std::string_view {"0.2.7-oracle"}; }
EOF

# 3. vendored oneTBB
if [ ! -f "$OUT/lib/libtbb.so.12" ]; then
    mkdir -p "$SCRATCH/tbb" && cd "$SCRATCH/tbb"
    cmake -G Ninja -DCMAKE_BUILD_TYPE=Release -DTBB_TEST=OFF -DTBB_STRICT=OFF "$REF/external/tbb" > cmake.log 2>&1
    ninja -j"$JOBS" tbb > build.log 2>&1
    cp -L "$(find . -name 'libtbb.so.12' | head -1)" "$OUT/lib/libtbb.so.12"
    ln -sf libtbb.so.12 "$OUT/lib/libtbb.so"
fi

# 4. vendored HYPRE (sequential + OpenMP)
if [ ! -f "$SCRATCH/hypre/libHYPRE.a" ]; then
    mkdir -p "$SCRATCH/hypre" && cd "$SCRATCH/hypre"
    cmake -G Ninja -DCMAKE_BUILD_TYPE=Release -DHYPRE_WITH_MPI=OFF -DHYPRE_ENABLE_MPI=OFF -DHYPRE_WITH_OPENMP=ON \
          -DHYPRE_ENABLE_OPENMP=ON -DHYPRE_BUILD_TESTS=OFF -DHYPRE_BUILD_EXAMPLES=OFF "$REF/external/hypre/src" > cmake.log 2>&1
    ninja -j"$JOBS" > build.log 2>&1
    [ -f libHYPRE.a ] || cp "$(find . -name 'libHYPRE.a' | head -1)" libHYPRE.a
fi

# 5. boost subset (boost/core/demangle.hpp)
if [ ! -d "$SCRATCH/boost" ]; then
    mkdir -p "$SCRATCH/boost" && tar -xzf "$REF/external/tecio/boost.tar.gz" -C "$SCRATCH/boost"
fi
BOOST_INC="$SCRATCH/boost"

# 6. drivers
HYPRE_INCS=""
for d in "$REF"/external/hypre/src/*/; do HYPRE_INCS="$HYPRE_INCS -I$d"; done
CXXFLAGS="-std=c++23 -O3 -DNDEBUG -fopenmp -DOPFLOW_WITH_OPENMP -DOPFLOW_WITH_HYPRE -DAMGCL_NO_BOOST -DSPDLOG_HEADER_ONLY \
 -DOPFLOW_PLATFORM_UNIX -DOPFLOW_HAS_MMAN_H -DOPFLOW_TEST_ENVIRONMENT -Wno-narrowing -w \
 -I$SCRATCH/shim -I$SCRATCH/hypre -I$SCRATCH/patched/include -I$SCRATCH/patched/src \
 -I$REF/external/spdlog/include -I$REF/external/tbb/include -I$REF/external/amgcl -I$REF/external/hypre/src $HYPRE_INCS \
 -I$REF/external/tecio/teciosrc -I$BOOST_INC"
cd "$HERE/ref_drivers"
if [ $# -gt 0 ]; then DRIVERS="$*"; else DRIVERS="$(ls *.cpp | sed 's/\.cpp$//')"; fi
pids=()
for drv in $DRIVERS; do
    src="$drv.cpp"; exe="$OUT/bin/$drv"
    if [ ! -f "$exe" ] || [ "$src" -nt "$exe" ]; then
        ( echo "build_ref: compiling $drv"; g++ $CXXFLAGS "$src" -o "$exe" "$SCRATCH/hypre/libHYPRE.a" -L"$OUT/lib" -ltbb -lm \
            -Wl,-rpath,'$ORIGIN/../lib' 2> "$SCRATCH/$drv.err" || { echo "build_ref: FAILED $drv (see $SCRATCH/$drv.err)"; tail -30 "$SCRATCH/$drv.err"; exit 1; } ) &
        pids+=($!)
        if [ "${#pids[@]}" -ge "$JOBS" ]; then wait "${pids[0]}"; pids=("${pids[@]:1}"); fi
    fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
echo "build_ref: done -> $OUT"
