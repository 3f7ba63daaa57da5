#!/bin/bash
# run here after tests/refresh_profiles_r02.sh: copies the raw material from gpurun_out/ into profiles/
set -e
cd /root/repo
O=gpurun_out
for w in c4a c1 c2 c3 c4b; do cp $O/r02_bench_$w.json profiles/r02_bench_$w.json; done
cp $O/r02_bench_c4a_reference_arm.json profiles/r02_bench_c4a_reference_arm.json
cp $O/r02_launches_c4a.csv profiles/r02_launches_c4a_final.csv
cp $O/r02_other_configs.txt profiles/r02_other_configs.txt
{
  echo "# ncu --set full --import-source on --clock-control none, \`python tests/perf_probe.py c4a\` (C4a: 1M paths, 16384^2, wide"
  echo "# coordinates), one launch of each of the three data-facing kernels (frame 3), B200, final build of round 2."
  echo "# Source: gpurun_out/r02_top3_c4a_final.ncu-rep (scratch, not tracked).  Launch list of a whole bench run:"
  echo "# r02_launches_c4a_final.csv.  \`traffic\` in bench.py's roofline object = dram read + write below (ncu_traffic.json)."
  echo
  python tests/ncu_summary.py $O/r02_top3_c4a_final.ncu-rep
} > profiles/r02_ncu_top3_c4a_final.txt
{
  echo "# ncu --set full, coverage mode AREA on C4a (SKB_COVERAGE_MODE=1 python tests/perf_probe.py c4a): k_area_bin (count"
  echo "# pass, write pass) and k_area_cover.  Source: gpurun_out/r02_area_c4a.ncu-rep (scratch)."
  echo
  python tests/ncu_summary.py $O/r02_area_c4a.ncu-rep
} > profiles/r02_ncu_area_c4a.txt
grep -E "Kernel Name|dram__bytes|gpu__time" profiles/r02_ncu_top3_c4a_final.txt
