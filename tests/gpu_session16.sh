#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -x -q -m gpu -k "area" > gpurun_out/s16_area.log 2>&1
tail -n 15 gpurun_out/s16_area.log
SKB_COVERAGE_MODE=1 timeout 300 python tests/perf_probe.py c1 p100k c4a 2>&1 | cut -c1-330
