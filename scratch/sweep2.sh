#!/bin/bash
run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['value'],1), round(d['roofline']['frac'],3))"; }
echo -n "window off: "; OPF_WINDOW=0 run
for cfg in "64 4 64" "32 8 64" "64 2 64" "64 4 16" "64 4 256" "32 4 64" "64 8 64" "128 2 64" "64 4 512"; do
  set -- $cfg
  echo -n "WTX=$1 WTY=$2 WCH=$3: "
  OPF_WTX=$1 OPF_WTY=$2 OPF_WCH=$3 run
done
echo -n "exact mode: "; run --mode exact
