#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/test_r2l.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|^E " gpurun_out/test_r2l.log | tail -12
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2l_drv.json 2> gpurun_out/bench_r2l_drv.err; echo "drv rc=$?"
python bench.py --steps 1000 --warmup 200 --skip-cpu > gpurun_out/bench_r2l.json 2> gpurun_out/bench_r2l.err; echo "long rc=$?"
PYTHONPATH=$PWD python tools/e2e_breakdown.py > gpurun_out/e2e_breakdown_r2l.txt 2>&1; cat gpurun_out/e2e_breakdown_r2l.txt
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r2l_ref.json 2> gpurun_out/bench_r2l_ref.err; echo "ref rc=$?"; cut -c1-250 gpurun_out/bench_r2l_ref.json
for e in AntMj AntMaze AntFlagrun; do python bench.py --env $e --steps 500 --warmup 200 --skip-cpu > gpurun_out/bench_r2l_$e.json 2> gpurun_out/bench_r2l_$e.err; echo "$e rc=$?"; cut -c1-200 gpurun_out/bench_r2l_$e.json; done
