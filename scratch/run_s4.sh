#!/bin/bash
python -m pytest tests/test_gpu_implicit.py tests/test_gpu_coefficients.py tests/test_gpu_frontend.py -q -x -k "not liddriven" 2>&1 | tail -3
python scratch/solve_prof.py 4097 | tail -1
python scratch/solve_prof.py 1025 | tail -1
OPF_MG_FUSED=0 python scratch/solve_prof.py 4097 | tail -1
python tools/bench_poisson_mgpu.py --size 257 --dim 3 2>&1 | tail -1 | cut -c1-250
python tools/bench_c5.py --size 1025 --nz 128 --steps 2 2>&1 | tail -1 | cut -c90-420
