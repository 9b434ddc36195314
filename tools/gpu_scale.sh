#!/bin/bash
# multi-GPU evidence on one box: tools/gpu_scale.sh N [tests]   -- bench.py (C2 weak scaling) and --config C5 at N GPUs; with "tests" the
# GPU test-suite first (the multi-GPU tests need >= 2 visible GPUs)
N=${1:-2}
mkdir -p gpurun_out
if [ "$2" = "tests" ]; then python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2_gputests_${N}gpu.txt; cat gpurun_out/r2_gputests_${N}gpu.txt; fi
for rep in ${OPF_SCALE_REPS:-a b}; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_scale_n${N}${rep}.out 2> gpurun_out/r2_scale_n${N}${rep}.err
grep '^{"metric' gpurun_out/r2_scale_n${N}${rep}.out > gpurun_out/r2_scale_n${N}${rep}.json; python - <<PY
import json
d=json.load(open("gpurun_out/r2_scale_n${N}${rep}.json"))
print("N=${N}${rep}", "GLUPS", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "kernel_ms", round(d["roofline"]["kernel_ms"],4), "parity_ok", d.get("parity_ok"), "e2e", round(d["e2e"]["value"],2), "batches", [round(x,2) for x in d["batches_ms"]], (d.get("comm_log_tail") or [""])[-1][-80:])
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --config C5 --steps 3 > gpurun_out/r2_c5_n${N}.out 2> gpurun_out/r2_c5_n${N}.err
grep '^{"metric' gpurun_out/r2_c5_n${N}.out > gpurun_out/r2_c5_n${N}.json; cut -c1-200 gpurun_out/r2_c5_n${N}.json; grep -o '"explicit_ms_per_step.*cell_steps_per_s": [0-9.]*' gpurun_out/r2_c5_n${N}.json; tail -3 gpurun_out/r2_c5_n${N}.err
