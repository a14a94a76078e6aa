#!/bin/bash
# round 2, GPU call AD: ncu of the raster-DEM target solve (C4), narrow bracket on and off
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:target_solve -s 1 -c 1 \
  -o gpurun_out/prof_solve_r2_c4 -f python scripts/e2e_breakdown.py c4 pinned 1 > gpurun_out/ncu_solve_r2_c4.log 2>&1
tail -2 gpurun_out/ncu_solve_r2_c4.log
I3B_NO_TIGHT_DEM=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:target_solve -s 1 -c 1 \
  -o gpurun_out/prof_solve_r2_c4_full -f python scripts/e2e_breakdown.py c4 pinned 1 > gpurun_out/ncu_solve_r2_c4_full.log 2>&1
tail -2 gpurun_out/ncu_solve_r2_c4_full.log
