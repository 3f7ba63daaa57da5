#!/bin/bash
cd "$GRAFT_REPO_ROOT"
{
for i in 1 2; do
echo "== long first (above median)"; timeout 300 python tests/perf_probe.py c2 c2clip c4bbatch64 2>&1 | cut -c1-200
echo "== SKB_WALK_NO_LONG"; SKB_WALK_NO_LONG=1 timeout 300 python tests/perf_probe.py c2 c2clip c4bbatch64 2>&1 | cut -c1-200
done
} > gpurun_out/s41.log 2>&1
cat gpurun_out/s41.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/s41_tests.log 2>&1
tail -n 2 gpurun_out/s41_tests.log
