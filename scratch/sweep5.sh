#!/bin/bash
run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['value'],1), round(d['roofline']['frac'],3))"; }
for st in 1 4 6; do for pd in 0 8; do echo -n "ST=$st PD=$pd: "; OPF_WST=$st OPF_WPD=$pd run; done; done
for cfg in "32 2 64" "64 2 64" "32 4 128" "32 4 32" "64 1 64"; do set -- $cfg; echo -n "ST=4 PD=0 WTX=$1 WTY=$2 WCH=$3: "; OPF_WST=4 OPF_WPD=0 OPF_WTX=$1 OPF_WTY=$2 OPF_WCH=$3 run; done
