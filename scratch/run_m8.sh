#!/bin/bash
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py 2>&1 | grep -E "MGPU|FAIL|Error" | head
for n in 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 100 --warmup 10 2>/dev/null | tail -1 > gpurun_out/scale_n$n.json
python -c "
import json; d=json.load(open('gpurun_out/scale_n$n.json')); print($n, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'])"
done
python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/scale_n1.json
python -c "
import json; d=json.load(open('gpurun_out/scale_n1.json')); print(1, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'])"
