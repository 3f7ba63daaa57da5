#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/s9_tests.log 2>&1
tail -n 3 gpurun_out/s9_tests.log
timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/s9_bench_c4a.json 2> gpurun_out/s9_bench_c4a.err
tail -c 800 gpurun_out/s9_bench_c4a.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/s9_bench_c4a.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','stages_ms')}, l['e2e']['ms_per_step'], l['e2e']['one_frame_at_a_time'], l.get('e2e_canvas'))
PY
timeout 600 python bench.py --workload c1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s9_bench_c1.json 2> gpurun_out/s9_bench_c1.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/s9_bench_c1.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','stages_ms')}, l['e2e']['ms_per_step'], l['e2e']['one_frame_at_a_time'], l.get('e2e_canvas'))
PY
