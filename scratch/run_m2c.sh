#!/bin/bash
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py > /tmp/mg.log 2>&1
grep -v "^\*\*\*\|OMP_NUM" /tmp/mg.log | grep -B2 -A12 "Traceback\|Error\|FAIL\|MGPU" | head -60
