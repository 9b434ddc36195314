#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/bench_configs.py 2>&1 | grep "C3\|C1'" | cut -c1-160
