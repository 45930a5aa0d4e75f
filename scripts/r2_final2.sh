#!/bin/bash
# round 2 closing pass on one B200 (v13 kernel: no halo warp, lighter ticket): GPU suite, smoke,
# both bench arms with the driver's defaults, ncu launch list, --set full captures summarised ON THE BOX (the .ncu-rep files stay in /tmp:
# gpurun copies back at most 64 MiB)
mkdir -p gpurun_out
t0=$(date +%s)
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2g_pytest.log; tail -3 gpurun_out/r2g_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2g_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/r2g_smoke.log
t1=$(date +%s)
timeout 1500 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench exit $? in $(( $(date +%s) - t1 )) s"
t1=$(date +%s)
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2g_bench_ref.json 2> gpurun_out/r2g_bench_ref.err; echo "reference arm exit $? in $(( $(date +%s) - t1 )) s"
cut -c1-400 gpurun_out/r2g_bench.json; cut -c1-300 gpurun_out/r2g_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 2 --warmup 1 --iters 50 --burnin 100 --init-iters 50 --batch-pairs 2 --batch-burnin 100 --e2e-its 50 --no-cpu > gpurun_out/r2g_ncu_launches.log 2>&1
R=/tmp/reps; mkdir -p $R
cap() {  # name, env, skip, args...
  local name=$1 envs=$2 skip=$3; shift 3
  env $envs timeout 900 python scripts/profile_target.py "$@" > gpurun_out/r2g_plain_$name.log 2>&1
  env $envs timeout 900 ncu --set full --clock-control none --import-source on -k regex:qgmap_iter -s $skip -c 1 -o $R/$name -f python scripts/profile_target.py "$@" > gpurun_out/r2g_ncu_$name.log 2>&1
  python scripts/ncu_summary.py $R/$name.ncu-rep gpurun_out/r2g_$name.txt "$(tail -1 gpurun_out/r2g_plain_$name.log)" > /dev/null 2>&1
  ncu -i $R/$name.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/r2g_${name}_source.csv.gz
}
cap 4k_v13 "QGMAP_X=0" 300 full 3 5 2160 3840 2 300 g
cap 640_v13_conv "QGMAP_X=0" 6000 full 3 5 480 640 2 6000 g
cap band270_v13 "QGMAP_X=0" 300 full 3 5 270 3840 2 300 g
cap c0_v13 "QGMAP_X=0" 1000 full 1 3 388 584 2 1000 g
cap super_v13 "QGMAP_X=0" 3000 super 3 5 480 640 2 3000 g
ls -la gpurun_out/r2g_*; echo "total $(( $(date +%s) - t0 )) s"
