#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
python tools/debug_wide.py 8 > gpurun_out/debug_wide8.log 2>&1; tail -20 gpurun_out/debug_wide8.log
python tools/warp_times.py AntGatherBulletEnv-v0 4096 > gpurun_out/warp_times_gather.log 2>&1; cat gpurun_out/warp_times_gather.log
python -m pytest tests -m gpu -x -q > gpurun_out/test_r2d.log 2>&1; tail -3 gpurun_out/test_r2d.log
