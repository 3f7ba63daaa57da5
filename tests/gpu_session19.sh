#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for v in base cw2m14 cw2m16 cw1m28 cw1m32 cw4m7 cnd32; do
echo "== $v"
SKB_LIB=gpurun_variants/$v.so timeout 300 python tests/perf_probe.py c1 c4a 2>&1 | tail -n 2 | cut -c1-200
done
