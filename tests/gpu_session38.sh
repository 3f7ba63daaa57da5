#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_walk$|k_cover|k_fine' --launch-skip 6 --launch-count 3 -o gpurun_out/r02_top3_c1_final -f python tests/perf_probe.py c1 > gpurun_out/s38.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'k_walk$|k_clip_rows|k_fine' --launch-skip 12 --launch-count 6 -o gpurun_out/r02_top_c2_final -f python tests/perf_probe.py c2clip >> gpurun_out/s38.log 2>&1
ls -la gpurun_out/r02_top3_c1_final.ncu-rep gpurun_out/r02_top_c2_final.ncu-rep
