#!/bin/bash
# round-2 evidence run A (1 GPU): GPU tests, bench with the configs block, C4 with the three Krylov-loop variants, launch list of one
# Poisson solve, ncu captures of the 2-D window skeleton (C1) and the WENO5 kernel (C3)
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/cond tools/cond_while_probe.cu -cudart shared && /tmp/cond
python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r2c_gputests.txt; cat gpurun_out/r2c_gputests.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err; tail -c 600 gpurun_out/r2c_bench_n1.json; tail -5 gpurun_out/r2c_bench_n1.err
for fk in 0 1 2; do OPF_FUSED_KRYLOV=$fk OPF_SOLVER_DEBUG=1 python bench.py --config C4 --no-cpu-baseline > gpurun_out/r2c_c4_fk$fk.json 2> gpurun_out/r2c_c4_fk$fk.err; cut -c1-330 gpurun_out/r2c_c4_fk$fk.json; tail -2 gpurun_out/r2c_c4_fk$fk.err | cut -c1-300; done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_c4_launches.csv python tools/solve_once.py > gpurun_out/r2c_solve_once.txt 2>&1
tail -3 gpurun_out/r2c_solve_once.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__inst_executed.sum
ncu --metrics $M --clock-control none -k regex:assign_kernel -s 8 -c 2 --csv --log-file gpurun_out/r2c_ncu_c3.csv python bench.py --config C3 --steps 10 --no-cpu-baseline > /dev/null 2>&1
python bench.py --config C3 --no-cpu-baseline | cut -c1-200
python bench.py --config C3B --no-cpu-baseline | cut -c1-200
