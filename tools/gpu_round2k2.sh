#!/bin/bash
# round 2k2: find the illegal access of fe_ld2d at 1025^2
cd /root/repo
mkdir -p gpurun_out/r2k2
O=gpurun_out/r2k2
for g in 1 0; do
  for w in 1 0; do
    OPF_GRAPHS=$g OPF_WINDOW=$w OPF_MG_COEF=0 OPF_MODE=fast timeout 120 tests/frontend/_bin/fe_ld2d --n 1025 --steps 2 --tol 1e-6 2>&1 | tail -1 | cut -c1-200 | sed "s/^/graphs=$g window=$w: /" | tee -a $O/ab.txt
  done
done
OPF_MG_COEF=0 OPF_MODE=fast timeout 600 compute-sanitizer --tool memcheck --print-limit 3 tests/frontend/_bin/fe_ld2d --n 1025 --steps 1 --tol 1e-3 > $O/sanitizer.txt 2>&1
grep -n "Invalid\|at .*kernel\|at opf\|by thread\|Address" $O/sanitizer.txt | head -20
timeout 600 python -m pytest tests/test_gpu_coefficients.py -x -q -m gpu 2>&1 | tail -3
