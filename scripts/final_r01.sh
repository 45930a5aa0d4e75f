# round-1 final measurement pass (one B200): tests, smoke, bench (both arms), band4k N=1, ncu launch list + full captures
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/final_pytest_gpu.log 2>&1; tail -3 gpurun_out/final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -3 gpurun_out/final_smoke.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -2 gpurun_out/final_bench.err; cut -c1-300 gpurun_out/final_bench.json
timeout 600 python bench.py --workload band4k --gpus 1 --steps 3 --warmup 1 --iters 100 --burnin 2000 > gpurun_out/final_band4k_n1.json 2> gpurun_out/final_band4k_n1.err; cut -c1-200 gpurun_out/final_band4k_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --burnin 100 --e2e-its 100 --no-cpu > gpurun_out/final_ncu_launches.log 2>&1
python scripts/profile_target.py full 2 9 480 640 4 6000 > gpurun_out/final_plain_full29.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qgmap_iter -s 6000 -c 1 -o gpurun_out/final_full29 -f python scripts/profile_target.py full 2 9 480 640 2 6000 > gpurun_out/final_ncu_full29.log 2>&1
python scripts/profile_target.py full 3 5 480 640 4 3000 > gpurun_out/final_plain_full35.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qgmap_iter -s 3000 -c 1 -o gpurun_out/final_full35 -f python scripts/profile_target.py full 3 5 480 640 2 3000 > gpurun_out/final_ncu_full35.log 2>&1
python scripts/quick_bench.py > gpurun_out/final_quick.log 2>&1; cat gpurun_out/final_quick.log
ls -la gpurun_out | tail -20
