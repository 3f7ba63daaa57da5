#!/bin/bash
cd "$GRAFT_REPO_ROOT"
SKB_COVERAGE_MODE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_area_cover' --launch-skip 1 -c 1 -f -o gpurun_out/r02_area2_c4a python tests/perf_probe.py c4a > gpurun_out/s17.log 2>&1
tail -n 2 gpurun_out/s17.log | cut -c1-200
