#!/bin/bash
N=$(nvidia-smi -L | wc -l); echo "gpus $N"
python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -3
for n in $N 2; do
[ "$n" -gt "$N" ] && continue
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 200 --warmup 10 2>/dev/null | tail -1 > gpurun_out/scale_r1f_n$n.json
python -c "
import json; d=json.load(open('gpurun_out/scale_r1f_n$n.json')); print($n, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'], d['clocks'])"
done
python bench.py --steps 200 --warmup 10 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/scale_r1f_n1.json
python -c "
import json; d=json.load(open('gpurun_out/scale_r1f_n1.json')); print(1, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'])"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 tools/bench_c5.py --size 1025 --nz 128 --steps 3 2>/dev/null | tail -1 | tee gpurun_out/c5_n$N.json | cut -c1-480
