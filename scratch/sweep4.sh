#!/bin/bash
run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['value'],1), round(d['roofline']['frac'],3))"; }
for pd in 0 1 2 4 8 16 32; do echo -n "PD=$pd: "; OPF_WPD=$pd run; done
for cfg in "32 4 64" "64 2 64" "32 2 256" "32 4 512"; do set -- $cfg; echo -n "PD=8 WTX=$1 WTY=$2 WCH=$3: "; OPF_WPD=8 OPF_WTX=$1 OPF_WTY=$2 OPF_WCH=$3 run; done
