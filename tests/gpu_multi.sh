#!/bin/bash
# usage: gpu_multi.sh N   (run under gpurun --gpus N)
cd "$GRAFT_REPO_ROOT"
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_c4a_n$N.json 2> gpurun_out/r02_bench_c4a_n$N.err
echo "rc=$?"
tail -c 600 gpurun_out/r02_bench_c4a_n$N.err
python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/r02_bench_c4a_n$N.json').read().strip().splitlines()[-1])
    print({k:l[k] for k in ('n_gpus','value','ms_per_step','stages_ms')}, l['e2e']['ms_per_step'], l['e2e']['one_frame_at_a_time'], l.get('gather'))
except Exception as e: print('ERR', e)
PY
