#!/bin/bash
# round-end evidence: GPU suite, bench line, launch list of the bench, ncu --set full of the dominant kernel, solver launch list
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; cut -c1-400 gpurun_out/bench_r1f.json
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_r1f_ref.json 2>> gpurun_out/bench_r1f.err; cut -c1-600 gpurun_out/bench_r1f_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/b_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tma_kernel -s 6 -c 1 -f -o gpurun_out/prof_ftcs3d_v7 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/b_ncu9.log 2>&1
python __graft_entry__.py smoke 2>&1 | tail -1
python tools/bench_configs.py > gpurun_out/bench_configs_r1f.jsonl 2>&1; cut -c1-150 gpurun_out/bench_configs_r1f.jsonl
