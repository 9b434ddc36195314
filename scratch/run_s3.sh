#!/bin/bash
for cs in 24 12 8 4; do echo "coarse sweeps $cs"; OPF_MG_COARSE_SWEEPS=$cs python scratch/solve_prof.py 1025 | tail -1; OPF_MG_COARSE_SWEEPS=$cs python scratch/solve_prof.py 4097 | tail -1;  OPF_MG_COARSE_SWEEPS=$cs python scratch/mgiters.py 2>&1 | grep "n=  257\|n= 1025" | grep "solver=5" ; done
