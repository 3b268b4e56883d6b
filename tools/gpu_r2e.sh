#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=$PWD
for l in 8 16; do python tools/debug_wide.py $l > gpurun_out/debug_wide$l.log 2>&1; grep -v "^   env" gpurun_out/debug_wide$l.log | tail -12; done
