#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for v in base w16 w20 w24 wb32 wb32m32 cw64 cm24; do
echo "== $v"
SKB_LIB=gpurun_variants/$v.so timeout 300 python tests/perf_probe.py c1 c4a 2>&1 | tail -n 2 | cut -c1-230
done > gpurun_out/s6_variants.log 2>&1
cat gpurun_out/s6_variants.log
