#!/bin/bash
python tools/bench_poisson_mgpu.py --size 4097 2>&1 | tail -1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/bench_poisson_mgpu.py --size 4097 2>&1 | tail -1
python tools/bench_poisson_mgpu.py --size 257 --dim 3 2>&1 | tail -1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/bench_poisson_mgpu.py --size 257 --dim 3 2>&1 | tail -1
