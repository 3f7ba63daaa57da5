#!/bin/bash
cd "$GRAFT_REPO_ROOT"
{
for v in cw160 cw192 cw160b28; do
echo "== $v"
SKB_LIB=gpurun_variants/$v.so timeout 300 python tests/perf_probe.py c1 c4a 2>&1 | tail -n 2 | cut -c1-200
done
} > gpurun_out/s28_variants.log 2>&1
cat gpurun_out/s28_variants.log
