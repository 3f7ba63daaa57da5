#!/bin/bash
cd "$GRAFT_REPO_ROOT"
{
echo "== long first (stride 2/4 only)"; timeout 300 python tests/perf_probe.py c1 c2 c2clip c3 c4bbatch64 2>&1 | cut -c1-200
} > gpurun_out/s40.log 2>&1
cat gpurun_out/s40.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/s40_tests.log 2>&1
tail -n 3 gpurun_out/s40_tests.log
timeout 600 python tests/gpu_fuzz.py 200 71000 2>&1 | tail -n 2
