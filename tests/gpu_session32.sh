#!/bin/bash
cd "$GRAFT_REPO_ROOT"
{
for v in r2; do
echo "== $v"
SKB_LIB=gpurun_variants/$v.so timeout 300 python tests/perf_probe.py c1 c2 c4a 2>&1 | tail -n 3 | cut -c1-200
done
SKB_LIB=gpurun_variants/r2.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "coverage or golden or fuzz or config" 2>&1 | tail -n 3
} > gpurun_out/s32_variants.log 2>&1
cat gpurun_out/s32_variants.log
