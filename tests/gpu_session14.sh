#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -x -q -m gpu -k "difference or golden_bit_exact" > gpurun_out/s14_diff.log 2>&1
tail -n 12 gpurun_out/s14_diff.log
