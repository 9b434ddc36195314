#!/bin/bash
# round 2k5: C5 momentum -- constant-bank vs array coefficient reads (two builds of fe_tg3d), skipped operator applications, launch list
cd /root/repo
mkdir -p gpurun_out/r2k5
O=gpurun_out/r2k5
for exe in fe_tg3d fe_tg3d_arrays; do
  [ -x tests/frontend/_bin/$exe ] || continue
  for skip in 1 0; do
    OPF_SOLVER_SKIP_E0=$skip OPF_GMRES_VERIFY=$((1-skip)) OPF_MODE=fast timeout 300 tests/frontend/_bin/$exe --n 1025 --nz 129 --steps 3 --tol 1e-8 2>&1 | tail -1 | cut -c90-420 | sed "s/^/$exe skip=$skip /" | tee -a $O/ab.txt
  done
done
timeout 900 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_implicit.py -x -q -m gpu > $O/tests.txt 2>&1
tail -4 $O/tests.txt
OPF_MODE=fast timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/tg3d_launches.csv tests/frontend/_bin/fe_tg3d --n 513 --nz 65 --steps 1 --tol 1e-8 > $O/ncu_run.txt 2>&1
python tools/launch_summary.py $O/tg3d_launches.csv 0 2>&1 | head -40 | tee $O/tg3d_launch_summary.txt
