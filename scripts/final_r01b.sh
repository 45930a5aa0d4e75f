# round-1 closing pass on one B200: GPU tests, smoke(), both bench arms
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/final3_pytest_gpu.log 2>&1; tail -3 gpurun_out/final3_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final3_smoke.log 2>&1; tail -3 gpurun_out/final3_smoke.log
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final3_bench_ref.json 2> gpurun_out/final3_bench_ref.err
timeout 600 python bench.py > gpurun_out/final3_bench.json 2> gpurun_out/final3_bench.err; tail -2 gpurun_out/final3_bench.err; cut -c1-400 gpurun_out/final3_bench.json
