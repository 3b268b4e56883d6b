#!/bin/bash
# A/B of the lane mappings: parity tests and bench for 4 / 8 / 16 lanes per env (HRL_B200_LANES)
L=${1:-ablanes}
mkdir -p gpurun_out
for lanes in 4 8 16; do
  HRL_B200_LANES=$lanes timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/test_${L}_l$lanes.log 2>&1
  echo "lanes=$lanes pytest rc=$?" | tee -a gpurun_out/test_${L}_l$lanes.log
  tail -4 gpurun_out/test_${L}_l$lanes.log
done
for lanes in 4 8 16; do
  for n in 4096 16384; do
    HRL_B200_LANES=$lanes timeout 600 python bench.py --steps 500 --warmup 200 --skip-cpu --envs-per-gpu $n > gpurun_out/bench_${L}_l${lanes}_n$n.json 2> gpurun_out/bench_${L}_l${lanes}_n$n.err
    echo "lanes=$lanes n=$n rc=$?"; cut -c1-260 gpurun_out/bench_${L}_l${lanes}_n$n.json
  done
done
