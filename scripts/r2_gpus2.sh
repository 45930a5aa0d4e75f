#!/bin/bash
# round 2, 2-GPU call (v13 kernel): multi-rank tests (peer-memory exchange fused into the iteration kernel, NCCL transport, band solver), bench N=2;
# super-pixel four-lane kernel at 6 resident CTAs (80 registers) against 5 (96)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2h2_smi.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_bands.py -x -q -m gpu > gpurun_out/r2h2_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2h2_pytest.log
tail -4 gpurun_out/r2h2_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --e2e-its 3000 --burnin 1000 --batch-burnin 1000 > gpurun_out/r2h2_bench_n2.json 2> gpurun_out/r2h2_bench_n2.err; echo "bench N=2 exit $?"
tail -3 gpurun_out/r2h2_bench_n2.err; cut -c1-700 gpurun_out/r2h2_bench_n2.json
C="super:480:640:3:5:3000:g,super:480:640:3:5:0:g,super:1080:1920:3:5:600:g"
python scripts/ab2.py m5 "$C" "m5=" > gpurun_out/r2h2_super_ab.txt 2>&1
QGMAP_LIB_PATH=build/libqgmap_sm6.so python scripts/ab2.py m6 "$C" "m6=" >> gpurun_out/r2h2_super_ab.txt 2>&1
cat gpurun_out/r2h2_super_ab.txt
