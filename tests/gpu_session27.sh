#!/bin/bash
cd "$GRAFT_REPO_ROOT"
{
for v in f12 f14 f16w4; do
echo "== $v"
SKB_LIB=gpurun_variants/$v.so timeout 300 python tests/perf_probe.py c1 c2 c3 c4a 2>&1 | tail -n 4 | cut -c1-200
done
for r in 10 12 13; do
echo "== frot resident $r"
SKB_WALK_RESIDENT=$r SKB_LIB=gpurun_variants/frot.so timeout 300 python tests/perf_probe.py c4a 2>&1 | tail -n 1 | cut -c1-200
done
} > gpurun_out/s27_variants.log 2>&1
cat gpurun_out/s27_variants.log
