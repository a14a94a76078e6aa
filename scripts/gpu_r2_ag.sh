#!/bin/bash
# round 2, GPU call AG: ncu of the raster-DEM target solve after the code diet
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:target_solve -s 1 -c 1 \
  -o gpurun_out/prof_solve_r2_c4_after -f python scripts/e2e_breakdown.py c4 pinned 1 > gpurun_out/ncu_solve_r2_c4_after.log 2>&1
tail -1 gpurun_out/ncu_solve_r2_c4_after.log
