#!/bin/bash
cd "$GRAFT_REPO_ROOT"
{
for v in wm15 wm16 wb32m30; do
echo "== $v"
SKB_LIB=gpurun_variants/$v.so timeout 300 python tests/perf_probe.py c1 c2clip c4a 2>&1 | tail -n 3 | cut -c1-200
done
} > gpurun_out/s29_variants.log 2>&1
cat gpurun_out/s29_variants.log
