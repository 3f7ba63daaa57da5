#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python bench.py > gpurun_out/r02_bench_c4a.json 2> gpurun_out/r02_bench_c4a.err
tail -c 600 gpurun_out/r02_bench_c4a.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r02_bench_c4a.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step')}, {k:v for k,v in l['e2e'].items() if k!='what'}, {k:v for k,v in l.get('e2e_canvas',{}).items() if k!='what'})
PY
timeout 600 python bench.py --workload c1 --no-cpu-baseline > gpurun_out/r02_bench_c1.json 2> gpurun_out/r02_bench_c1.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r02_bench_c1.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step')}, {k:v for k,v in l['e2e'].items() if k!='what'}, {k:v for k,v in l.get('e2e_canvas',{}).items() if k!='what'})
PY
