#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for m in 0 1; do
echo "== SKB_WALK_MODE=$m"
SKB_WALK_MODE=$m timeout 300 python tests/perf_probe.py c1 c2 c4a 2>&1 | tail -4
done > gpurun_out/s4_walkmodes.log 2>&1
tail -n 12 gpurun_out/s4_walkmodes.log
