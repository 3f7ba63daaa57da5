#!/bin/bash
cd "$GRAFT_REPO_ROOT"
{
for v in base sinksm base sinksm; do
echo "== $v"; SKB_LIB=gpurun_variants/$v.so timeout 300 python tests/perf_probe.py c1 c4a 2>&1 | tail -n 2 | cut -c1-200
done
} > gpurun_out/s36.log 2>&1
cat gpurun_out/s36.log
