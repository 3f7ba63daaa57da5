#!/bin/bash
cd "$GRAFT_REPO_ROOT"
SKB_COVERAGE_MODE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 50 --csv --log-file gpurun_out/s3_area_launch_c4a.csv python tests/perf_probe.py c4a > gpurun_out/s3_a.log 2>&1
SKB_COVERAGE_MODE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_area_cover|k_area_bin' -c 3 -f -o gpurun_out/r02_area_c4a python tests/perf_probe.py c4a > gpurun_out/s3_b.log 2>&1
tail -3 gpurun_out/s3_a.log gpurun_out/s3_b.log
