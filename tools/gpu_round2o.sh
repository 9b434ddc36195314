#!/bin/bash
# round 2o: launch geometry / occupancy of the direct skeleton on the C5 momentum operators (fe_tg3d, 1024 x 1024 x 128 cells)
cd /root/repo
mkdir -p gpurun_out/r2o
O=gpurun_out/r2o
run() { exe=$1; shift; env "$@" OPF_MODE=fast timeout 300 tests/frontend/_bin/$exe --n 1025 --nz 129 --steps 3 --tol 1e-8 2>&1 | tail -1 | cut -c108-300 | sed "s/^/$exe $* : /" | tee -a $O/ab.txt; }
run fe_tg3d OPF_X=0
run fe_tg3d OPF_TY=2
run fe_tg3d OPF_TY=2 OPF_CH=16
run fe_tg3d OPF_TY=1 OPF_CH=16
run fe_tg3d OPF_TX=64 OPF_TY=4 OPF_CH=16
if [ -x tests/frontend/_bin/fe_tg3d_occ ]; then
run fe_tg3d_occ OPF_TY=2
run fe_tg3d_occ OPF_TY=2 OPF_CH=16
run fe_tg3d_occ OPF_TX=64 OPF_TY=4 OPF_CH=16
fi
