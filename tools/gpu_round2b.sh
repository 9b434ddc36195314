#!/bin/bash
# round-2 evidence run B (2 GPUs): the multi-GPU tests, bench N=2 with the halo time line and without
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -12 > gpurun_out/r2c_gputests_multi.txt; cat gpurun_out/r2c_gputests_multi.txt
OPF_HALO_DEBUG=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err
grep '^{"metric' gpurun_out/r2c_bench_n2.json | cut -c1-400; grep "opf halo" gpurun_out/r2c_bench_n2.err | tail -6
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c_bench_n2b.json 2> gpurun_out/r2c_bench_n2b.err
grep '^{"metric' gpurun_out/r2c_bench_n2b.json | cut -c1-300; tail -3 gpurun_out/r2c_bench_n2b.err
