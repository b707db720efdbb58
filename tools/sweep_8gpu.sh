#!/bin/bash
# 1/2/4/8-GPU sweeps of BASELINE configs 5, 4, 1 and 2 on ONE 8-GPU box (weak scaling: fixed work per GPU).
# N = 1, 2 and 4 run concurrently on disjoint GPUs (0 | 1-2 | 3-6), then N = 8.  Output: gpurun_out/sweep_c<config>_n<N>.json
run() { n=$1; devs=$2; port=$3; shift 3
  CUDA_VISIBLE_DEVICES=$devs python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@"; }
sweep() { c=$1; shift
  (run 1 0 29601 --config $c "$@" > gpurun_out/sweep_c${c}_n1.json 2> gpurun_out/sweep_c${c}_n1.err) &
  (run 2 1,2 29602 --config $c "$@" > gpurun_out/sweep_c${c}_n2.json 2> gpurun_out/sweep_c${c}_n2.err) &
  (run 4 3,4,5,6 29603 --config $c "$@" > gpurun_out/sweep_c${c}_n4.json 2> gpurun_out/sweep_c${c}_n4.err) &
  wait
  run 8 0,1,2,3,4,5,6,7 29604 --config $c "$@" > gpurun_out/sweep_c${c}_n8.json 2> gpurun_out/sweep_c${c}_n8.err; }
python -c "import torch" 2>/dev/null      # page the image in once
sweep 5 --steps 1 --warmup 3 --no-e2e
sweep 4 --steps 5 --warmup 3
sweep 1 --steps 20 --warmup 10
sweep 2 --steps 5 --warmup 3 --no-cpu-baseline --no-extras
for f in gpurun_out/sweep_c*_n*.json; do echo "== $f"; head -c 300 $f; echo; done
