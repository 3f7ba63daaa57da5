#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/s12_tests.log 2>&1
tail -n 3 gpurun_out/s12_tests.log
timeout 600 python tests/e2e_probe.py 3 2>&1 | tail -n 12
