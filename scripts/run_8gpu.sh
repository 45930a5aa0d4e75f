for tr in p2p nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29501 bench.py --workload band4k --gpus 8 --steps 3 --warmup 1 --iters 100 --burnin 2000 --band-transport $tr > gpurun_out/band4k_n8_$tr.json 2> gpurun_out/band4k_n8_$tr.err
tail -1 gpurun_out/band4k_n8_$tr.err; cut -c1-160 gpurun_out/band4k_n8_$tr.json
done
