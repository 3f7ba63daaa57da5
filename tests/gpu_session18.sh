#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for v in a32m12 a32m14 a32m16 a32w2m24 a32w8m6; do
echo "== $v"
SKB_COVERAGE_MODE=1 SKB_LIB=gpurun_variants/$v.so timeout 300 python tests/perf_probe.py c1 c4a 2>&1 | tail -n 2 | cut -c1-200
done
