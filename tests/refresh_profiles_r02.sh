#!/bin/bash
# run on the GPU box (gpurun): regenerates the raw material of profiles/r02_* into gpurun_out/
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python bench.py > $O/r02_bench_c4a.json 2> $O/r02_bench_c4a.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_c4a_reference_arm.json 2> $O/r02_bench_ref.err
for w in c1 c2 c3 c4b; do
  timeout 600 python bench.py --workload $w --steps 6 --warmup 3 --no-cpu-baseline > $O/r02_bench_$w.json 2> $O/r02_bench_$w.err
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_c4a.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-canvas-e2e > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_walk$|k_cover|k_fine' --launch-skip 6 --launch-count 3 -o $O/r02_top3_c4a_final -f python tests/perf_probe.py c4a > $O/r02_ncu_top3.log 2>&1
timeout 900 python tests/perf_probe.py c1 p100k c4a c2 c2clip c3 c4bbatch64 > $O/r02_other_configs.txt 2>&1
{ echo "== SKB_COVERAGE_MODE=1 (AREA)"; SKB_COVERAGE_MODE=1 timeout 600 python tests/perf_probe.py c1 c4a; echo "== SKB_WALK_MODE=1 (row-parallel walk)"; SKB_WALK_MODE=1 timeout 600 python tests/perf_probe.py c1 c4a; } >> $O/r02_other_configs.txt 2>&1
cut -c1-420 $O/r02_bench_c4a.json; cut -c1-300 $O/r02_bench_c4a_reference_arm.json; cut -c1-160 $O/r02_other_configs.txt
