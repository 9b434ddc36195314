#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 "$@" > /tmp/o.txt 2>&1; tail -1 /tmp/o.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['value'],1), round(d['roofline']['frac'],3))" 2>/dev/null || tail -3 /tmp/o.txt; }
for cx in 2 4; do for by in 4 8; do for ch in 64 128; do echo -n "TMA CX=$cx BY=$by CH=$ch: "; OPF_TCX=$cx OPF_TBY=$by OPF_TCH=$ch run; done; done; done
echo -n "exact: "; run --mode exact
