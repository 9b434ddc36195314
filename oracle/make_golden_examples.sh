#!/usr/bin/env bash
# oracle/make_golden_examples.sh -- TEST INFRASTRUCTURE.  Regenerates tests/golden/examples.json and
# tests/golden/liddriven2d_n65_s1000.json from the UNMODIFIED reference (needs /root/reference and a finished oracle/build_ref.sh):
#   * examples.json: the computation of examples/FTCS2D/FTCS-OMP.cpp:8-31 (1025^2, 5000 steps, u = 0, Dirichlet 1) through
#     oracle/_ref/bin/ref_explicit (the example itself needs HDF5 for its writer, which this image lacks), centre value u[512,512];
#   * liddriven2d_n65_s1000.json: examples/LidDriven/LidDriven2D.cpp compiled UNCHANGED with the build_ref.sh flags (g++ -std=c++23,
#     vendored HYPRE + oneTBB), run to completion (1000 steps, ~6 min on 8 cores); last Tecplot zone of u.tec / v.tec / p.tec.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${OPF_REFERENCE:-/root/reference}"
SCRATCH="${OPF_REF_SCRATCH:-/tmp/opflow_ref_build}"
OUT="$HERE/_ref"
WORK="$(mktemp -d)"
export LD_LIBRARY_PATH="$OUT/lib:${LD_LIBRARY_PATH:-}"

"$OUT/bin/ref_explicit" --case ftcs2d --n 1025 --steps 5000 --threads "$(nproc)" --dump "$WORK/ftcs.opfd"
python - "$WORK/ftcs.opfd" "$ROOT" <<'PY'
import json, sys
sys.path.insert(0, sys.argv[2])
from oracle import oracle as O
s, e, a = O.read_opfd(sys.argv[1])
json.dump({"ftcs_omp_1025_5000_center": float(a[512 - s[0], 512 - s[1]]),
           "_source": "unmodified reference: oracle/_ref/bin/ref_explicit --case ftcs2d --n 1025 --steps 5000 (the computation of examples/FTCS2D/FTCS-OMP.cpp:8-31), u[512,512]"},
          open(sys.argv[2] + "/tests/golden/examples.json", "w"), indent=1)
PY

HYPRE_INCS=""
for d in "$REF"/external/hypre/src/*/; do HYPRE_INCS="$HYPRE_INCS -I$d"; done
/usr/bin/g++ -std=c++23 -O3 -DNDEBUG -fopenmp -DOPFLOW_WITH_OPENMP -DOPFLOW_WITH_HYPRE -DAMGCL_NO_BOOST -DSPDLOG_HEADER_ONLY \
    -DOPFLOW_PLATFORM_UNIX -DOPFLOW_HAS_MMAN_H -DOPFLOW_TEST_ENVIRONMENT -Wno-narrowing -w \
    -I"$SCRATCH/shim" -I"$SCRATCH/hypre" -I"$SCRATCH/patched/include" -I"$SCRATCH/patched/src" \
    -I"$REF/external/spdlog/include" -I"$REF/external/tbb/include" -I"$REF/external/amgcl" -I"$REF/external/hypre/src" $HYPRE_INCS \
    -I"$REF/external/tecio/teciosrc" -I"$SCRATCH/boost" \
    "$REF/examples/LidDriven/LidDriven2D.cpp" -o "$WORK/lid_ref" "$SCRATCH/hypre/libHYPRE.a" -L"$OUT/lib" -ltbb -lm -Wl,-rpath,"$OUT/lib"
(cd "$WORK" && ./lid_ref > run.log 2>&1)
python - "$WORK" "$ROOT" <<'PY'
import json, re, sys
import numpy as np
work, root = sys.argv[1], sys.argv[2]
out = {"_source": "unmodified reference: examples/LidDriven/LidDriven2D.cpp (n=65, 1000 steps) built with g++ + vendored HYPRE/TBB by the oracle/build_ref.sh recipe; last Tecplot zone of u.tec / v.tec / p.tec (10 printed digits)"}
for name in "uvp":
    last = open(f"{work}/{name}.tec").read().rsplit("ZONE\n", 1)[1].splitlines()
    m = re.match(r"I = (\d+) J = (\d+)", last[1])
    I, J = int(m.group(1)), int(m.group(2))
    vals = np.array([float(x) for x in last[3:3 + 3 * I * J]])
    out[name] = {"I": I, "J": J, "time": float(last[2].split("=")[1]), "values": vals[2 * I * J:].tolist()}
json.dump(out, open(root + "/tests/golden/liddriven2d_n65_s1000.json", "w"))
PY
rm -rf "$WORK"
echo "examples.json and liddriven2d_n65_s1000.json written"

# lid-driven cavity at sizes where the momentum operators need a multigrid preconditioner (nu dt / h^2 = 0.8 at 129^2 nodes, 3.3 at 257^2):
# oracle/ref_drivers/ref_ld2d.cpp run by the unmodified reference, every 2nd / 4th value of u, v, p kept
"$OUT/bin/ref_ld2d" --n 129 --steps 3 --threads 8 --tol 1e-10 --stride 2 --dump "$ROOT/tests/golden/ld2d_n129_s3" > /dev/null
"$OUT/bin/ref_ld2d" --n 257 --steps 3 --threads 8 --tol 1e-10 --stride 4 --dump "$ROOT/tests/golden/ld2d_n257_s3" > /dev/null
echo "ld2d_n129_s3_{u,v,p}.opfd and ld2d_n257_s3_{u,v,p}.opfd written"
