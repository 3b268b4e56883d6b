#!/bin/bash
# One GPU visit with the evidence the judge reads: parity tests (verbose log), the driver-protocol bench line, the ncu
# launch list of the same command and one full capture of the dominant kernel.  Usage (under gpurun): bash tools/gpu_round.sh LABEL
L=${1:-rX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s -rs > gpurun_out/tests_$L.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/tests_$L.log; grep -E "passed|failed" gpurun_out/tests_$L.log | tail -2
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$L.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$L.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${L}_drv.json 2> gpurun_out/bench_${L}_drv.err; echo "bench (driver protocol) rc=$?"
python bench.py --steps 1000 --warmup 200 --skip-cpu > gpurun_out/bench_${L}_long.json 2> gpurun_out/bench_${L}_long.err; echo "bench (1000 steps) rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_${L}_ref.json 2>/dev/null; echo "reference arm rc=$?"
if [ -z "$NO_NCU" ]; then
  # (a number printed under ncu is never a bench value: both logs go to gpurun_out/ncu_*.log)
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$L.csv \
    python bench.py --steps 20 --warmup 5 --skip-cpu --no-graph > gpurun_out/ncu_list_$L.log 2>&1; echo "ncu list rc=$?"
  ncu --set full --import-source on --clock-control none --kernel-name ant_env_kernel --launch-skip 230 --launch-count 1 \
    -f -o gpurun_out/prof_$L python bench.py --steps 20 --warmup 5 --skip-cpu --no-graph > gpurun_out/ncu_full_$L.log 2>&1; echo "ncu full rc=$?"
fi
