#!/bin/bash
python -m pytest tests/test_gpu_hostpipe.py -q -x 2>&1 | tail -8
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 6 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'])"
