#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for v in fdiv frot; do
echo "== $v"
SKB_LIB=gpurun_variants/$v.so timeout 300 python tests/perf_probe.py c1 c2 c3 c4a 2>&1 | tail -n 4 | cut -c1-200
done > gpurun_out/s26_variants.log 2>&1
cat gpurun_out/s26_variants.log
