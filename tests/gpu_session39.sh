#!/bin/bash
cd "$GRAFT_REPO_ROOT"
{
echo "== long first"; timeout 300 python tests/perf_probe.py c1 c2 c2clip c3 c4bbatch64 2>&1 | cut -c1-200
echo "== SKB_WALK_NO_LONG"; SKB_WALK_NO_LONG=1 timeout 300 python tests/perf_probe.py c1 c2 c2clip c3 c4bbatch64 2>&1 | cut -c1-200
} > gpurun_out/s39.log 2>&1
cat gpurun_out/s39.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/s39_tests.log 2>&1
tail -n 3 gpurun_out/s39_tests.log
