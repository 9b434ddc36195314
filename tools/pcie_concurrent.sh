#!/bin/bash
# tools/pcie_concurrent.sh N -- N processes, one per GPU, each moving 1.08 GB up and 1.08 GB down concurrently (tools/pcie_both.cu)
N=${1:-1}
nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o /tmp/pcie_both tools/pcie_both.cu -cudart shared || exit 1
for n in 1 $N; do
  echo "== $n process(es)"
  for ((g = 0; g < n; ++g)); do /tmp/pcie_both $g 10 & done
  wait
done
