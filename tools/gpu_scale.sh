#!/bin/bash
# multi-GPU bench lines under the driver's protocol: bash tools/gpu_scale.sh N LABEL
N=$1; L=${2:-r2}
mkdir -p gpurun_out
run() {  # env steps warmup tag
  if [ "$N" = 1 ]; then python bench.py --gpus 1 --env $1 --steps $2 --warmup $3 --skip-cpu > gpurun_out/scale_${L}_$1_n${N}_$4.json 2> gpurun_out/scale_${L}_$1_n${N}_$4.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --env $1 --steps $2 --warmup $3 --skip-cpu > gpurun_out/scale_${L}_$1_n${N}_$4.json 2> gpurun_out/scale_${L}_$1_n${N}_$4.err; fi
  echo "$1 N=$N steps=$2 rc=$?"; python -c "
import json,sys
d=json.load(open('gpurun_out/scale_${L}_$1_n${N}_$4.json'))
print('   value %.4g  ms/step %.4f  b2b %.4g  e2e %.4g  clocks %s' % (d['value'], d['ms_per_step'], d['config']['value_back_to_back'], d['e2e']['value'], d['clocks']))
print('   per-rank mean us', d['config']['per_rank_step_us']['mean'], 'max', d['config']['per_rank_step_us']['max'], 'slowest', d['config']['per_rank_step_us']['slowest_rank'])"
}
run AntGather 20 5 drv
run AntMaze 20 5 drv
run AntFlagrun 20 5 drv
if [ -z "$SHORT" ]; then
  run AntGather 20 5 drv2
  run AntGather 1000 200 long
  run AntMaze 500 200 long
  run AntFlagrun 500 200 long
fi
