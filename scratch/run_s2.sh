#!/bin/bash
python -m pytest tests/test_gpu_implicit.py tests/test_gpu_explicit.py -q -x 2>&1 | tail -3
python scratch/solve_prof.py 4097
python scratch/solve_prof.py 1025
