#!/bin/bash
# round 2k: coefficient-carrying multigrid (LidDriven2D at 129^2 / 257^2 / 1025^2), CSR export, full GPU suite, default bench, C5
cd /root/repo
mkdir -p gpurun_out/r2k
O=gpurun_out/r2k
nvidia-smi --query-gpu=name --format=csv,noheader > $O/gpu.txt
timeout 900 python -m pytest tests/test_gpu_frontend.py -x -q -m gpu -k "lid_driven" > $O/ld_tests.txt 2>&1
tail -5 $O/ld_tests.txt
for n in 257 513 1025; do
  for opt in 0 1; do
    OPF_MG_COEF=$opt OPF_MODE=fast OPF_SOLVER_DEBUG=0 timeout 600 tests/frontend/_bin/fe_ld2d --n $n --steps 5 --tol 1e-10 2>&1 | tail -1 | sed "s/^/fe n=$n MG_COEF=$opt /" | tee -a $O/ld2d_times.txt
  done
  timeout 900 oracle/_ref/bin/ref_ld2d --n $n --steps 2 --threads 16 --tol 1e-10 2>&1 | tail -1 | sed "s/^/ref n=$n 16 threads /" | tee -a $O/ld2d_times.txt
done
OPF_MODE=fast OPF_SOLVER_DEBUG=1 timeout 300 tests/frontend/_bin/fe_ld2d --n 1025 --steps 2 --tol 1e-10 > $O/ld2d_debug.txt 2>&1
timeout 2400 python -m pytest tests -x -q -m gpu > $O/gputests.txt 2>&1
tail -5 $O/gputests.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
cat $O/bench_n1.json | cut -c1-600
timeout 900 python bench.py --config C5 --steps 3 --warmup 1 > $O/c5_n1.json 2> $O/c5_n1.err
cat $O/c5_n1.json | cut -c1-900
