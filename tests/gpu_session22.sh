#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for v in both both32; do
for cap in 0 10 14 18 22 26 30; do
echo "== $v cap $cap"
SKB_WALK_SMEM_EDGES=$cap SKB_LIB=gpurun_variants/$v.so timeout 300 python tests/perf_probe.py c4a 2>&1 | tail -n 1 | cut -c1-200
done
done > gpurun_out/s22_variants.log 2>&1
cat gpurun_out/s22_variants.log
