#!/bin/bash
# round 2, 2-GPU call: multi-rank tests (peer-memory exchange fused into the iteration kernel, NCCL transport, band solver), smoke, bench N=2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2g2_smi.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_bands.py -x -q -m gpu > gpurun_out/r2g2_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2g2_pytest.log
tail -8 gpurun_out/r2g2_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2g2_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/r2g2_smoke.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --e2e-its 6000 --burnin 1000 --batch-burnin 1000 > gpurun_out/r2g2_bench_n2.json 2> gpurun_out/r2g2_bench_n2.err; echo "bench N=2 exit $?"
tail -3 gpurun_out/r2g2_bench_n2.err; cat gpurun_out/r2g2_bench_n2.json | cut -c1-1500
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 2 --e2e-its 600 --burnin 1000 --batch-pairs 0 --band-transport nccl > gpurun_out/r2g2_bench_n2_nccl.json 2> gpurun_out/r2g2_bench_n2_nccl.err; echo "bench N=2 nccl exit $?"
cat gpurun_out/r2g2_bench_n2_nccl.json | cut -c1-400
