#!/bin/bash
# A/B of the solver sweeps in one GPU visit: the default build (velocity-space sweep) against the same sources built with
# -DHRL_DELASSUS=1 (Delassus-space fast path; build it first:
#   python -c "from hrl_pybullet_envs_b200 import _cabi; _cabi.build(force=True, defines={'HRL_DELASSUS': 1}, out='hrl_pybullet_envs_b200/libhrl_b200_ds.so')").
# The parity tests run against the Delassus build (HRL_B200_LIB), the default build is covered by every other visit.
# Usage (under gpurun): bash tools/gpu_ab_sweep.sh LABEL [ENV ...]
L=${1:-ab}; shift
ENVS=${@:-AntGather}
mkdir -p gpurun_out
HRL_B200_LIB=$PWD/hrl_pybullet_envs_b200/libhrl_b200_ds.so python -m pytest tests -m gpu -q -x -rs > gpurun_out/tests_$L.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|Error" gpurun_out/tests_$L.log | tail -3
for E in $ENVS; do
  for V in vel delassus; do
    LIB=""; [ $V = delassus ] && LIB=$PWD/hrl_pybullet_envs_b200/libhrl_b200_ds.so
    HRL_B200_LIB=$LIB python bench.py --env $E --steps 300 --warmup 20 --skip-cpu > gpurun_out/bench_${L}_${E}_$V.json 2> gpurun_out/bench_${L}_${E}_$V.err
    echo "$E $V rc=$? $(python -c "import json,sys; d=json.loads(open('gpurun_out/bench_${L}_${E}_$V.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['config'].get('ms_per_step_cuda_graph'))")"
  done
done
