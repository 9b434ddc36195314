#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-1500
