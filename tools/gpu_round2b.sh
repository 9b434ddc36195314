#!/bin/bash
# round-2 evidence run B (2 GPUs): all GPU tests (multi-GPU ones included), C4 after the multigrid changes + launch list, bench N=2 with the halo time line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r2d_gputests.txt; cat gpurun_out/r2d_gputests.txt
for fk in 1 2; do OPF_FUSED_KRYLOV=$fk python bench.py --config C4 --no-cpu-baseline > gpurun_out/r2d_c4_fk$fk.json 2> gpurun_out/r2d_c4_fk$fk.err; cut -c1-330 gpurun_out/r2d_c4_fk$fk.json; tail -2 gpurun_out/r2d_c4_fk$fk.err | cut -c1-300; done
OPF_MG_FAST_XFER=0 python bench.py --config C4 --no-cpu-baseline | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d_c4_launches.csv python tools/solve_once.py > gpurun_out/r2d_solve_once.txt 2>&1
tail -2 gpurun_out/r2d_solve_once.txt
OPF_HALO_DEBUG=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err
grep '^{"metric' gpurun_out/r2d_bench_n2.json | cut -c1-400; grep "opf halo" gpurun_out/r2d_bench_n2.err | tail -6
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2d_bench_n2b.json 2> gpurun_out/r2d_bench_n2b.err
grep '^{"metric' gpurun_out/r2d_bench_n2b.json | cut -c1-300; tail -3 gpurun_out/r2d_bench_n2b.err
