#!/bin/bash
python tools/bench_c5.py --size 257 --nz 64 --steps 3 2>&1 | tail -2 | cut -c1-500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/bench_c5.py --size 257 --nz 64 --steps 3 2>&1 | tail -2 | cut -c1-500
python tools/bench_c5.py --size 1025 --nz 128 --steps 3 2>&1 | tail -1 | cut -c1-500
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/bench_c5.py --size 1025 --nz 128 --steps 3 2>&1 | tail -1 | cut -c1-500
