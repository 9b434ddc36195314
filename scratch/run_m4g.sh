#!/bin/bash
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py > /tmp/mg.log 2>&1
grep "MGPU\|FAIL\|iters\|even split" /tmp/mg.log | head -12
