#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -x -q -m gpu -k "difference or golden_bit_exact" > gpurun_out/s8_diff_tests.log 2>&1
tail -n 15 gpurun_out/s8_diff_tests.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s8_tests.log 2>&1
tail -n 3 gpurun_out/s8_tests.log
