#!/bin/bash
# geometry sweep for the FTCS3D kernel
for cfg in "128 4 32" "128 4 8" "128 4 128" "64 8 32" "32 16 32" "128 2 32" "128 1 32" "64 4 32" "256 2 32" "128 4 512"; do
  set -- $cfg
  echo -n "TX=$1 TY=$2 CH=$3: "
  OPF_TX=$1 OPF_TY=$2 OPF_CH=$3 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['frac'])"
done
