#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_walk$|k_cover' --launch-skip 4 --launch-count 2 -o gpurun_out/r02b_walk_cover_c4a -f python tests/perf_probe.py c4a > gpurun_out/s24.log 2>&1
tail -n 2 gpurun_out/s24.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
