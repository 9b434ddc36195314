#!/bin/bash
# round-2 evidence run A: GPU tests (2 GPUs: multi-GPU tests included), bench N=1 with the configs block, bench N=2 with the halo time line,
# ncu captures of the 2-D window skeleton (C1) and the WENO5 kernel (C3)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r2b_gputests.txt; cat gpurun_out/r2b_gputests.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; tail -c 3000 gpurun_out/r2b_bench_n1.json; tail -5 gpurun_out/r2b_bench_n1.err
OPF_HALO_DEBUG=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err
grep '^{"metric' gpurun_out/r2b_bench_n2.json | cut -c1-400; grep "opf halo" gpurun_out/r2b_bench_n2.err | tail -6
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2b_bench_n2b.json 2> gpurun_out/r2b_bench_n2b.err
grep '^{"metric' gpurun_out/r2b_bench_n2b.json | cut -c1-300
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__inst_executed.sum
ncu --metrics $M --clock-control none -k regex:assign_kernel -s 8 -c 2 --csv --log-file gpurun_out/r2b_ncu_c3.csv python bench.py --config C3 --steps 10 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:window_kernel -s 30 -c 2 --csv --log-file gpurun_out/r2b_ncu_c1.csv python bench.py --config C1 --steps 10 --no-cpu-baseline > /dev/null 2>&1
tail -3 gpurun_out/r2b_ncu_c3.csv | cut -c1-400
