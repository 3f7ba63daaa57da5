#!/bin/bash
# run here after refresh_profiles.sh: copies the raw material from gpurun_out/ into profiles/
set -e
cd /root/repo
cp gpurun_out/bench_c1.json profiles/r01_bench_c1.json
cp gpurun_out/bench_ref.json profiles/r01_bench_c1_reference_arm.json
cp gpurun_out/launches.csv profiles/r01_launches_c1_final.csv
cp gpurun_out/other_configs.txt profiles/r01_other_configs.txt
{
  echo "# ncu --set full --import-source on --clock-control none, \`python tests/perf_probe.py c1\` (C1: 10k paths, 4096^2),"
  echo "# one launch of each of the three data-facing kernels (frame 3), B200.  Source: gpurun_out/top3_c1.ncu-rep"
  echo "# (scratch, not tracked); per-section details in r01_ncu_top3_c1_details.txt; launch list of a whole bench run in"
  echo "# r01_launches_c1_final.csv.  \`traffic\` in bench.py's roofline object = dram read + write below."
  echo
  python tests/ncu_summary.py gpurun_out/top3_c1.ncu-rep
} > profiles/r01_ncu_top3_c1.txt
ncu -i gpurun_out/top3_c1.ncu-rep --page details > profiles/r01_ncu_top3_c1_details.txt 2>/dev/null
grep -E "Kernel Name|dram__bytes" profiles/r01_ncu_top3_c1.txt
