#!/bin/bash
# round-2 first visit: parity tests, driver-protocol bench (20 steps), long bench, other envs
L=${1:-r2a}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/test_$L.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/test_$L.log
tail -5 gpurun_out/test_$L.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$L.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$L.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${L}_drv.json 2> gpurun_out/bench_${L}_drv.err
echo "bench drv rc=$?"; cut -c1-300 gpurun_out/bench_${L}_drv.json
python bench.py --steps 1000 --warmup 200 --skip-cpu > gpurun_out/bench_$L.json 2> gpurun_out/bench_$L.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_$L.json
for e in AntMj AntMaze AntFlagrun; do
  python bench.py --env $e --steps 20 --warmup 5 --skip-cpu > gpurun_out/bench_${L}_$e.json 2> gpurun_out/bench_${L}_$e.err
  echo "$e rc=$?"; cut -c1-200 gpurun_out/bench_${L}_$e.json
done
