#!/bin/bash
# round-2 GPU session 2: AREA mode parity + first timings
cd "$GRAFT_REPO_ROOT"
python -m pytest tests -m gpu -x -q -k "area" -s > gpurun_out/s2_area_tests.log 2>&1; tail -15 gpurun_out/s2_area_tests.log
python -m pytest tests -m gpu -x -q -k "not area" > gpurun_out/s2_tests.log 2>&1; tail -3 gpurun_out/s2_tests.log
SKB_COVERAGE_MODE=1 timeout 300 python tests/perf_probe.py c1 p100k c4a > gpurun_out/s2_area_probe.log 2>&1; cat gpurun_out/s2_area_probe.log
