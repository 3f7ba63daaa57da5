#!/bin/bash
cd "$GRAFT_REPO_ROOT"
{
echo "== default"; timeout 300 python tests/perf_probe.py c1 c2clip c4a 2>&1 | tail -n 3 | cut -c1-200
echo "== spec"; SKB_LIB=gpurun_variants/spec.so timeout 300 python tests/perf_probe.py c1 c2clip c4a 2>&1 | tail -n 3 | cut -c1-200
} > gpurun_out/s34_variants.log 2>&1
cat gpurun_out/s34_variants.log
