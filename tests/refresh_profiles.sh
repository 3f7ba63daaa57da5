#!/bin/bash
# run on the GPU box: regenerates the raw material of profiles/ into gpurun_out/
python bench.py > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_walk$|k_cover|k_fine' --launch-skip 6 --launch-count 3 -o gpurun_out/top3_c1 -f python tests/perf_probe.py c1 > gpurun_out/ncu_top3.log 2>&1
python tests/perf_probe.py c1 p100k c4a c2 c2clip c3 c4bbatch64 > gpurun_out/other_configs.txt 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_c1.json | cut -c1-400; cut -c1-150 gpurun_out/other_configs.txt; cat gpurun_out/bench_ref.json | cut -c1-500
