#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/s15_tests.log 2>&1
tail -n 3 gpurun_out/s15_tests.log
timeout 300 python tests/perf_probe.py c1 c4a 2>&1 | cut -c1-200
