#!/bin/bash
cd "$GRAFT_REPO_ROOT"
{
for c in -2 0 25 100; do
echo "== carveout $c"; SKB_WALK_CARVEOUT=$c timeout 300 python tests/perf_probe.py c4a 2>&1 | tail -n 1 | cut -c1-200
done
} > gpurun_out/s35.log 2>&1
cat gpurun_out/s35.log
