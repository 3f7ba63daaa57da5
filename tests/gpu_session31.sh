#!/bin/bash
cd "$GRAFT_REPO_ROOT"
{
for v in in2u2 in1u2 in2u3 in2u4 in3u3 in4u4; do
echo "== $v"
SKB_LIB=gpurun_variants/$v.so timeout 300 python tests/perf_probe.py c1 c2 c4a 2>&1 | tail -n 3 | cut -c1-200
done
SKB_LIB=gpurun_variants/in2u2.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "coverage or golden or fuzz or config" 2>&1 | tail -n 3
} > gpurun_out/s31_variants.log 2>&1
cat gpurun_out/s31_variants.log
