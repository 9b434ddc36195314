#!/bin/bash
# round-2 evidence run I (1 GPU): full suite, default bench line, C4, C5, ncu launch list of the bench step, ncu --set full of the TMA kernel,
# ncu metrics of the multigrid / Krylov kernels of one Poisson solve
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2i_gputests.txt; cat gpurun_out/r2i_gputests.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err; tail -c 300 gpurun_out/r2i_bench_n1.json; tail -3 gpurun_out/r2i_bench_n1.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2i_bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/r2i_bench_reference.json
python bench.py --config C4 > gpurun_out/r2i_c4.json 2> gpurun_out/r2i_c4.err; cut -c1-300 gpurun_out/r2i_c4.json
python bench.py --config C1 > gpurun_out/r2i_c1.json 2> gpurun_out/r2i_c1.err; cut -c1-700 gpurun_out/r2i_c1.json
python bench.py --config C5 --steps 3 > gpurun_out/r2i_c5_n1.json 2> gpurun_out/r2i_c5_n1.err; cut -c1-1500 gpurun_out/r2i_c5_n1.json; tail -3 gpurun_out/r2i_c5_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2i_bench_launches.csv python bench.py --steps 3 --warmup 3 --batches 1 --no-configs --no-cpu-baseline --e2e-steps 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tma_kernel -s 6 -c 1 -o gpurun_out/r2i_tma_full python bench.py --steps 3 --warmup 3 --batches 1 --no-configs --no-cpu-baseline --e2e-steps 1 > /dev/null 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct
ncu --metrics $M --clock-control none -s 3700 -c 260 --csv --log-file gpurun_out/r2i_c4_kernels.csv python tools/solve_once.py > gpurun_out/r2i_solve_once.txt 2>&1
tail -2 gpurun_out/r2i_solve_once.txt; ls -la gpurun_out | tail -12
