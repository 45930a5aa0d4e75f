timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu > gpurun_out/pytest_multirank.log 2>&1; tail -15 gpurun_out/pytest_multirank.log
for tr in p2p nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501 bench.py --workload band4k --gpus 2 --steps 3 --warmup 1 --iters 100 --burnin 500 --band-transport $tr > gpurun_out/band4k_n2_$tr.json 2> gpurun_out/band4k_n2_$tr.err
tail -2 gpurun_out/band4k_n2_$tr.err; cut -c1-160 gpurun_out/band4k_n2_$tr.json
done
