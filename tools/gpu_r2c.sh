#!/bin/bash
L=r2c
mkdir -p gpurun_out
python tools/debug_wide.py 8 > gpurun_out/debug_wide8.log 2>&1; tail -20 gpurun_out/debug_wide8.log
for lanes in 4 8 16; do
  HRL_B200_LANES=$lanes ncu --set full --import-source on --clock-control none --kernel-name regex:ant_env --launch-skip 230 --launch-count 1 \
    -f -o gpurun_out/prof_${L}_l$lanes python bench.py --steps 20 --warmup 5 --skip-cpu --no-graph > gpurun_out/ncu_${L}_l$lanes.log 2>&1
  echo "ncu lanes=$lanes rc=$?"
done
ls -la gpurun_out/*.ncu-rep
