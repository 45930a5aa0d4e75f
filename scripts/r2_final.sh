#!/bin/bash
# round 2 closing pass on one B200: full GPU suite, smoke, both bench arms (defaults = what the driver runs), ncu launch list + --set full captures
mkdir -p gpurun_out
t0=$(date +%s)
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2f_pytest.log; tail -3 gpurun_out/r2f_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2f_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/r2f_smoke.log
t1=$(date +%s)
timeout 1500 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench exit $? in $(( $(date +%s) - t1 )) s"
t1=$(date +%s)
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; echo "reference arm exit $? in $(( $(date +%s) - t1 )) s"
cut -c1-400 gpurun_out/r2f_bench.json; cut -c1-300 gpurun_out/r2f_bench_ref.json
# ncu launch list of a short bench run (every launch inside the timed regions is qgmap_iter_kernel<5,false,false>)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 1 --iters 50 --burnin 100 --init-iters 50 --batch-pairs 2 --batch-burnin 100 --e2e-its 50 --no-cpu > gpurun_out/r2f_ncu_launches.log 2>&1
# --set full: 4K grey-level pair after 300 iterations (the bench's regime: wide beliefs, fp16 one-sector layout); same with the fp32 layout;
# 480x640 L=3 K=5 after 6000 iterations (converged: tap cache path); one band of eight (270 x 3840)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qgmap_iter -s 300 -c 1 -o gpurun_out/r2f_4k_v12 -f python scripts/profile_target.py full 3 5 2160 3840 2 300 g > gpurun_out/r2f_ncu_4k.log 2>&1
QGMAP_TAPS=f32 timeout 900 ncu --set full --clock-control none -k regex:qgmap_iter -s 300 -c 1 -o gpurun_out/r2f_4k_v12_f32taps -f python scripts/profile_target.py full 3 5 2160 3840 2 300 g > gpurun_out/r2f_ncu_4k_f32.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:qgmap_iter -s 6000 -c 1 -o gpurun_out/r2f_640_v12_conv -f python scripts/profile_target.py full 3 5 480 640 2 6000 g > gpurun_out/r2f_ncu_640.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:qgmap_iter -s 300 -c 1 -o gpurun_out/r2f_band270_v12 -f python scripts/profile_target.py full 3 5 270 3840 2 300 g > gpurun_out/r2f_ncu_band.log 2>&1
tail -1 gpurun_out/r2f_ncu_4k.log gpurun_out/r2f_ncu_4k_f32.log gpurun_out/r2f_ncu_640.log gpurun_out/r2f_ncu_band.log
ls -la gpurun_out/r2f_*.ncu-rep; echo "total $(( $(date +%s) - t0 )) s"
