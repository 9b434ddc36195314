#!/bin/bash
# round 2k3: after the guard-band fix -- ld2d tests + timings (1025^2 too), CSR export tests, full GPU suite, default bench, C5
cd /root/repo
mkdir -p gpurun_out/r2k3
O=gpurun_out/r2k3
timeout 900 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_coefficients.py -x -q -m gpu -k "lid_driven or csr" > $O/ld_tests.txt 2>&1
tail -5 $O/ld_tests.txt
for n in 1025 2049; do
  for opt in 0 1; do
    OPF_MG_COEF=$opt OPF_MODE=fast timeout 600 tests/frontend/_bin/fe_ld2d --n $n --steps 5 --tol 1e-10 2>&1 | tail -1 | sed "s/^/fe n=$n MG_COEF=$opt /" | tee -a $O/ld2d_times.txt
  done
done
OPF_MODE=fast compute-sanitizer --tool memcheck --print-limit 3 tests/frontend/_bin/fe_ld2d --n 1025 --steps 1 --tol 1e-3 > $O/sanitizer.txt 2>&1
tail -3 $O/sanitizer.txt
timeout 2400 python -m pytest tests -x -q -m gpu > $O/gputests.txt 2>&1
tail -5 $O/gputests.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
cat $O/bench_n1.json | cut -c1-400
timeout 900 python bench.py --config C5 --steps 3 --warmup 1 > $O/c5_n1.json 2> $O/c5_n1.err
cat $O/c5_n1.json | cut -c1-1200
