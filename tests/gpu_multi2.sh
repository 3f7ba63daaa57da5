#!/bin/bash
# usage: gpu_multi2.sh N [workload]  (run under gpurun --gpus N): GPU tests of the new pieces, then the banded bench
cd "$GRAFT_REPO_ROOT"
N=$1; W=${2:-c4a}
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "culled or band" > gpurun_out/m2_tests.log 2>&1; tail -n 2 gpurun_out/m2_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $W --steps 5 --warmup 3 > gpurun_out/r02b_bench_${W}_n$N.json 2> gpurun_out/r02b_bench_${W}_n$N.err
echo "rc=$?"
tail -c 1500 gpurun_out/r02b_bench_${W}_n$N.err
python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/r02b_bench_${W}_n$N.json').read().strip().splitlines()[-1])
    print({k:l[k] for k in ('n_gpus','value','ms_per_step','stages_ms')}); print(l['e2e']); print(l.get('gather'), l.get('band_display_lists'))
except Exception as e: print('ERR', e)
PY
