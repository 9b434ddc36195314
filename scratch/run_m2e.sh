#!/bin/bash
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py > /tmp/mg.log 2>&1
grep "MGPU\|FAIL\|iters" /tmp/mg.log | head
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/bench_poisson_mgpu.py --size 257 --dim 3 2>&1 | tail -1 | cut -c1-300
