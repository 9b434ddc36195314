#!/bin/bash
# full state check: GPU tests, bench, launch list, ncu of the dominant kernel
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; tail -c 3000 gpurun_out/bench_r1c.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/b_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tma_kernel -s 6 -c 1 -f -o gpurun_out/prof_ftcs3d_v5 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/b_ncu7.log 2>&1
python __graft_entry__.py smoke 2>&1 | tail -2
