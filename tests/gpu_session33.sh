#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 python tests/perf_probe.py c1 c2clip c3 c4a c4bbatch64 2>&1 | cut -c1-250 | tee gpurun_out/s33_probe.log
SKB_COVERAGE_MODE=1 timeout 300 python tests/perf_probe.py c4a 2>&1 | cut -c1-250 | tee -a gpurun_out/s33_probe.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/s33_tests.log 2>&1
tail -n 3 gpurun_out/s33_tests.log
timeout 600 python tests/gpu_fuzz.py 200 61000 2>&1 | tail -n 2
