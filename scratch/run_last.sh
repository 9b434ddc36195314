#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 10 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(2, d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 tools/bench_c5.py --size 1025 --nz 128 --steps 3 2>/dev/null | tail -1 | cut -c80-420
