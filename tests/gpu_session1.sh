#!/bin/bash
# round-2 GPU session 1: tests, walk occupancy sweep on C4a, ncu captures of C2 (clips) and C3
cd "$GRAFT_REPO_ROOT"
python -m pytest tests -m gpu -x -q > gpurun_out/s1_tests.log 2>&1; tail -3 gpurun_out/s1_tests.log
for s in 0 24576 32768 49152 65536 110000; do
  echo "== SKB_WALK_SMEM=$s" >> gpurun_out/s1_walk_smem.log
  SKB_WALK_SMEM=$s timeout 300 python tests/perf_probe.py c4a >> gpurun_out/s1_walk_smem.log 2>&1
done
cat gpurun_out/s1_walk_smem.log | grep -o "SKB_WALK_SMEM=.*\|'walk': [0-9.]*"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_walk|k_cover|k_fine|k_clip_rows' -c 4 -f -o gpurun_out/r02_top_c2 python tests/perf_probe.py c2clip > gpurun_out/s1_ncu_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_walk|k_cover|k_fine|k_blur_h|k_blur_v' -c 7 -f -o gpurun_out/r02_top_c3 python tests/perf_probe.py c3 > gpurun_out/s1_ncu_c3.log 2>&1
tail -2 gpurun_out/s1_ncu_c2.log gpurun_out/s1_ncu_c3.log
