#!/bin/bash
# round-2 evidence run E (1 GPU): all GPU tests after the new device nodes / solver changes, C4 timing, launch list of one solve
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2e_gputests.txt; cat gpurun_out/r2e_gputests.txt
for fk in 1 2; do OPF_FUSED_KRYLOV=$fk python bench.py --config C4 --no-cpu-baseline > gpurun_out/r2e_c4_fk$fk.json 2> gpurun_out/r2e_c4_fk$fk.err; cut -c1-330 gpurun_out/r2e_c4_fk$fk.json; tail -2 gpurun_out/r2e_c4_fk$fk.err | cut -c1-300; done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2e_c4_launches.csv python tools/solve_once.py > gpurun_out/r2e_solve_once.txt 2>&1
tail -2 gpurun_out/r2e_solve_once.txt
