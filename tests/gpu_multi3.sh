#!/bin/bash
cd "$GRAFT_REPO_ROOT"
N=$1
for mode in rank owner rank owner; do
SKB_BENCH_SHM_TOUCH=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/m3_$mode.json 2> gpurun_out/m3_$mode.err
python - <<PY
import json
l=json.loads(open('gpurun_out/m3_$mode.json').read().strip().splitlines()[-1])
print('$mode', l['ms_per_step'], 'e2e', l['e2e']['ms_per_step'], 'serial', l['e2e']['one_frame_at_a_time']['ms_per_step'], 'via0', l['e2e']['via_rank0_canvas']['ms_per_step'], l['e2e'].get('rank0_bound_to_numa_node'))
PY
done
