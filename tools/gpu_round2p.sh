#!/bin/bash
# round 2p: final 1-GPU validation -- full GPU suite, default bench, C5, LidDriven2D timings, smoke
cd /root/repo
mkdir -p gpurun_out/r2r
O=gpurun_out/r2r
timeout 2400 python -m pytest tests -q -m gpu > $O/gputests.txt 2>&1
tail -4 $O/gputests.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
cut -c1-260 $O/bench_n1.json
timeout 900 python bench.py --config C5 --steps 3 --warmup 1 > $O/c5_n1.json 2> $O/c5_n1.err
grep -o '"explicit_ms_per_step.*cell_steps_per_s": [0-9.]*' $O/c5_n1.json
for n in 1025 2049; do
  OPF_MODE=fast timeout 600 tests/frontend/_bin/fe_ld2d --n $n --steps 5 --tol 1e-10 2>&1 | tail -1 | cut -c60-330 | sed "s/^/fe_ld2d n=$n /" | tee -a $O/ld2d_times.txt
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
