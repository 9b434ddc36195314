#!/bin/bash
# round 2k4: C5 momentum A/B -- constant-bank coefficients (OPF_UNIFORM) and coefficient hierarchy (OPF_MG_COEF)
cd /root/repo
mkdir -p gpurun_out/r2k4
O=gpurun_out/r2k4
for u in 1 0; do for c in 1 0; do
  OPF_UNIFORM=$u OPF_MG_COEF=$c OPF_MODE=fast timeout 300 tests/frontend/_bin/fe_tg3d --n 1025 --nz 129 --steps 3 --tol 1e-8 2>&1 | tail -1 | cut -c90-420 | sed "s/^/uniform=$u mg_coef=$c /" | tee -a $O/ab.txt
done; done
OPF_SOLVER_DEBUG=1 OPF_MODE=fast timeout 300 tests/frontend/_bin/fe_tg3d --n 513 --nz 65 --steps 1 --tol 1e-8 2>&1 | grep opf_solver | head -8 | cut -c1-260
