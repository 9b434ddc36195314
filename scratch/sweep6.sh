#!/bin/bash
run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['value'],1), round(d['roofline']['frac'],3))"; }
echo -n "TMA off (window ST=1): "; OPF_TMA=0 OPF_WST=1 run
for by in 4 8; do for ch in 32 64 128 512; do echo -n "TMA BY=$by CH=$ch: "; OPF_TBY=$by OPF_TCH=$ch run; done; done
echo -n "exact TMA: "; run --mode exact
