#!/bin/bash
# round 2m: ncu --set full of the semi-implicit momentum operators (direct skeleton) of fe_tg3d; final 1-GPU validation
cd /root/repo
mkdir -p gpurun_out/r2m
O=gpurun_out/r2m
OPF_MODE=fast timeout 900 ncu --set full --clock-control none --kernel-name-base demangled --kernel-name 'regex:assign_kernel.*Div' --launch-skip 4 --launch-count 3 \
    -f -o /tmp/momentum tests/frontend/_bin/fe_tg3d --n 513 --nz 65 --steps 0 --tol 1e-8 > $O/ncu_run.txt 2>&1
ncu -i /tmp/momentum.ncu-rep --page raw --csv > $O/momentum_raw.csv 2> $O/ncu_export.err
ls -la /tmp/momentum.ncu-rep $O/momentum_raw.csv
timeout 2400 python -m pytest tests -x -q -m gpu > $O/gputests.txt 2>&1
tail -4 $O/gputests.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
cut -c1-300 $O/bench_n1.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
