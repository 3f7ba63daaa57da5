#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s5_tests.log 2>&1
tail -n 3 gpurun_out/s5_tests.log
for R in 0 3 5 7 9 11; do
echo "== SKB_WALK_RESIDENT=$R"
SKB_WALK_RESIDENT=$R timeout 300 python tests/perf_probe.py c4a 2>&1 | tail -n 1 | cut -c1-330
done > gpurun_out/s5_walk_resident.log 2>&1
cat gpurun_out/s5_walk_resident.log
timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/s5_bench_c4a.json 2> gpurun_out/s5_bench_c4a.err
tail -c 1500 gpurun_out/s5_bench_c4a.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/s5_bench_c4a.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','stages_ms')}, l['e2e'], l.get('e2e_canvas'))
PY
