#!/bin/bash
# usage: scale_run.sh N   -- runs the default (pairs) bench and the band4k bench on N GPUs, outputs under gpurun_out/
N=$1
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --no-cpu > gpurun_out/scale_pairs_n1.json 2> gpurun_out/scale_pairs_n1.err
  python bench.py --workload band4k --gpus 1 --steps 3 --warmup 1 --iters 100 --burnin 2000 > gpurun_out/scale_band4k_n1.json 2> gpurun_out/scale_band4k_n1.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus $N --no-cpu > gpurun_out/scale_pairs_n$N.json 2> gpurun_out/scale_pairs_n$N.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29501 bench.py --workload band4k --gpus $N --steps 3 --warmup 1 --iters 100 --burnin 2000 > gpurun_out/scale_band4k_n$N.json 2> gpurun_out/scale_band4k_n$N.err
fi
