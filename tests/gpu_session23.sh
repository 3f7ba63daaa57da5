#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/s23_tests.log 2>&1
tail -n 3 gpurun_out/s23_tests.log
timeout 300 python tests/perf_probe.py c1 c2clip c3 c4a 2>&1 | cut -c1-330 | tee gpurun_out/s23_probe.log
