#!/bin/bash
# what the driver runs at round end, on one box: GPU tests, smoke, the bench's own arm and the reference arm
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/final_bench.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/final_ref.json
