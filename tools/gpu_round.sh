#!/bin/bash
# One GPU visit: parity tests, bench of the default build + tuning variants, ncu launch list and
# one full capture of the dominant kernel.  Usage (under gpurun): bash tools/gpu_round.sh LABEL [variants...]
L=${1:-rX}; shift
VARIANTS="$@"
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/test_$L.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/test_$L.log
tail -3 gpurun_out/test_$L.log
python bench.py --steps 1000 --warmup 200 --skip-cpu > gpurun_out/bench_$L.json 2> gpurun_out/bench_$L.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_$L.json
for v in $VARIANTS; do
  HRL_B200_LIB=$PWD/hrl_pybullet_envs_b200/libhrl_b200_$v.so python bench.py --steps 1000 --warmup 200 --skip-cpu > gpurun_out/bench_${L}_$v.json 2> gpurun_out/bench_${L}_$v.err
  echo "variant $v rc=$?"; cut -c1-200 gpurun_out/bench_${L}_$v.json
done
if [ -z "$NO_NCU" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$L.csv \
    python bench.py --steps 20 --warmup 200 --skip-cpu > gpurun_out/ncu_list_$L.log 2>&1
  ncu --set full --import-source on --clock-control none --kernel-name regex:ant_env --launch-skip 215 --launch-count 1 \
    -f -o gpurun_out/prof_$L python bench.py --steps 20 --warmup 200 --skip-cpu > gpurun_out/ncu_full_$L.log 2>&1
  echo "ncu rc=$?"
fi
ls -la gpurun_out | tail -20
