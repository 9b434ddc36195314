#!/bin/bash
cd /root/repo
mkdir -p gpurun_out/r2q
run() { env "$@" OPF_MODE=fast timeout 300 tests/frontend/_bin/fe_tg3d --n 1025 --nz 129 --steps 3 --tol 1e-8 2>&1 | tail -1 | cut -c108-260 | sed "s/^/$* : /" | tee -a gpurun_out/r2q/ab.txt; }
run OPF_X=0
run OPF_X=0
run OPF_TY=2
OPF_MODE=fast timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2q/tg3d_launches.csv tests/frontend/_bin/fe_tg3d --n 513 --nz 65 --steps 1 --tol 1e-8 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2q/tg3d_launches.csv 0 14 2>&1 | cut -c1-200
