#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/s11_tests.log 2>&1
tail -n 3 gpurun_out/s11_tests.log
timeout 600 python tests/perf_probe.py c1 c2 c3 c4a 2>&1 | cut -c1-330 > gpurun_out/s11_probe.log
cat gpurun_out/s11_probe.log
for v in v14 v34 v11; do
echo "== $v"
SKB_LIB=gpurun_variants/$v.so timeout 300 python tests/perf_probe.py c1 c4a 2>&1 | tail -n 2 | cut -c1-230
done > gpurun_out/s11_variants.log 2>&1
cat gpurun_out/s11_variants.log
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -k regex:'k_walk$' -c 3 --csv --log-file gpurun_out/s11_walk_lanes.csv python tests/perf_probe.py c4a > /dev/null 2>&1
grep -v "^==" gpurun_out/s11_walk_lanes.csv | cut -d, -f5,13- | tail -n 9
