#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rs -s > gpurun_out/test_r2j_l4.log 2>&1; echo "lanes=4 rc=$?"; grep -E "passed|failed|FAILED|Error" gpurun_out/test_r2j_l4.log | tail -15
for lanes in 8 16; do
  HRL_B200_LANES=$lanes python -m pytest tests -m gpu -q > gpurun_out/test_r2j_l$lanes.log 2>&1; echo "lanes=$lanes rc=$?"; grep -E "passed|failed|FAILED" gpurun_out/test_r2j_l$lanes.log | tail -8
done
