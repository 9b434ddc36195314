#!/bin/bash
for tx in 32 64 128; do for ch in 16 32 64 128; do for pd in 0 8; do OPF_WTX=$tx OPF_WCH=$ch OPF_WPD=$pd python scratch/w2d.py 2>&1 | tail -1; done; done; done
