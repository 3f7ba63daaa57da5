#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/s10_tests.log 2>&1
tail -n 3 gpurun_out/s10_tests.log
timeout 600 python tests/perf_probe.py c2 c2clip c3 2>&1 | cut -c1-330 > gpurun_out/s10_probe.log
cat gpurun_out/s10_probe.log
