#!/bin/bash
# round 2l: red-black Gauss-Seidel smoother, affine fused residual; full suite; C4 with both smoothers; C5; ld2d timings
cd /root/repo
mkdir -p gpurun_out/r2l
O=gpurun_out/r2l
timeout 900 python -m pytest tests/test_gpu_implicit.py -x -q -m gpu -s -k "red_black" > $O/rb_tests.txt 2>&1
tail -5 $O/rb_tests.txt; grep "iterations:" $O/rb_tests.txt
for r in 1 2; do
  OPF_C4_RELAX=$r timeout 600 python bench.py --config C4 --steps 10 --warmup 3 2> $O/c4_relax$r.err | tee $O/c4_relax$r.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C4 relax=$r', d.get('ms_per_step'), {k:v for k,v in d.items() if 'iter' in k or 'rel' in k})"
done
for n in 1025 2049; do
  OPF_MODE=fast timeout 600 tests/frontend/_bin/fe_ld2d --n $n --steps 5 --tol 1e-10 2>&1 | tail -1 | sed "s/^/fe n=$n /" | tee -a $O/ld2d_times.txt
done
timeout 2400 python -m pytest tests -x -q -m gpu > $O/gputests.txt 2>&1
tail -5 $O/gputests.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
cat $O/bench_n1.json | cut -c1-300
timeout 900 python bench.py --config C5 --steps 3 --warmup 1 > $O/c5_n1.json 2> $O/c5_n1.err
cat $O/c5_n1.json | cut -c600-1100
