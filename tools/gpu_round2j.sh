#!/bin/bash
# round-2 evidence run J (1 GPU): C1 / C4 after the skeleton-gating fix, per-kernel ncu metrics of one Poisson solve, launch list of a solve
mkdir -p gpurun_out
python bench.py --config C1 --no-cpu-baseline > gpurun_out/r2j_c1.json 2> gpurun_out/r2j_c1.err; cut -c1-500 gpurun_out/r2j_c1.json
python bench.py --config C4 --no-cpu-baseline > gpurun_out/r2j_c4.json 2> gpurun_out/r2j_c4.err; cut -c1-300 gpurun_out/r2j_c4.json
python bench.py --config C3B --no-cpu-baseline > gpurun_out/r2j_c3b.json 2> gpurun_out/r2j_c3b.err; cut -c1-200 gpurun_out/r2j_c3b.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,lts__t_sector_hit_rate.pct
ncu --metrics $M --clock-control none -s 3500 -c 300 --csv --log-file gpurun_out/r2j_c4_kernels.csv python tools/solve_once.py > gpurun_out/r2j_solve_once.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2j_c4_launches.csv python tools/solve_once.py > gpurun_out/r2j_solve_once2.txt 2>&1
tail -2 gpurun_out/r2j_solve_once2.txt
