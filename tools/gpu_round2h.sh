#!/bin/bash
# round-2 evidence run H (1 GPU): full GPU suite, smoke, default bench line (configs block incl. C1 graph replay), C5 at reduced and full per-GPU size
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2h_gputests.txt; cat gpurun_out/r2h_gputests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err; tail -c 400 gpurun_out/r2h_bench_n1.json; tail -3 gpurun_out/r2h_bench_n1.err
OPF_C5_N=257 OPF_C5_NZ=64 python bench.py --config C5 --steps 2 > gpurun_out/r2h_c5_small.json 2> gpurun_out/r2h_c5_small.err; cut -c1-900 gpurun_out/r2h_c5_small.json; tail -3 gpurun_out/r2h_c5_small.err
OPF_SOLVER_DEBUG=1 python bench.py --config C5 --steps 2 > gpurun_out/r2h_c5_n1.json 2> gpurun_out/r2h_c5_n1.err; cut -c1-1200 gpurun_out/r2h_c5_n1.json; tail -8 gpurun_out/r2h_c5_n1.err | cut -c1-300
