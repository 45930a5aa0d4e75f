#!/bin/bash
# round 2, 8-GPU call: bench at N=8 (default settings = what the driver runs) and N=4
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2g8_smi.log 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 > gpurun_out/r2g8_bench_n8.json 2> gpurun_out/r2g8_bench_n8.err; echo "bench N=8 exit $?"
tail -3 gpurun_out/r2g8_bench_n8.err; cut -c1-300 gpurun_out/r2g8_bench_n8.json
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 4 --e2e-its 9000 > gpurun_out/r2g8_bench_n4.json 2> gpurun_out/r2g8_bench_n4.err; echo "bench N=4 exit $?"
cut -c1-300 gpurun_out/r2g8_bench_n4.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 8 --e2e-its 600 --batch-pairs 0 --band-transport nccl > gpurun_out/r2g8_bench_n8_nccl.json 2> gpurun_out/r2g8_bench_n8_nccl.err; echo "bench N=8 nccl exit $?"
cut -c1-300 gpurun_out/r2g8_bench_n8_nccl.json
