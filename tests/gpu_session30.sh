#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/s30_tests.log 2>&1
tail -n 3 gpurun_out/s30_tests.log
timeout 900 python tests/gpu_fuzz.py 300 52000 > gpurun_out/s30_fuzz.log 2>&1; tail -n 3 gpurun_out/s30_fuzz.log
timeout 300 python tests/perf_probe.py c1 c4a 2>&1 | cut -c1-250 | tee gpurun_out/s30_probe.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
