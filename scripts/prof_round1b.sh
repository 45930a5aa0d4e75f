set -x
mkdir -p gpurun_out
python scripts/profile_target.py full 3 5 480 640 4 3000 > gpurun_out/p2_plain_full35.log 2>&1
python scripts/profile_target.py full 1 3 388 584 4 1000 > gpurun_out/p2_plain_full13.log 2>&1
python scripts/profile_target.py super 3 5 480 640 4 1500 > gpurun_out/p2_plain_super35.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:qgmap_iter -s 3000 -c 1 -o gpurun_out/p2_full35 -f python scripts/profile_target.py full 3 5 480 640 2 3000 > gpurun_out/p2_ncu_full35.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:qgmap_iter -s 1000 -c 1 -o gpurun_out/p2_full13 -f python scripts/profile_target.py full 1 3 388 584 2 1000 > gpurun_out/p2_ncu_full13.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:qgmap_iter -s 1500 -c 1 -o gpurun_out/p2_super35 -f python scripts/profile_target.py super 3 5 480 640 2 1500 > gpurun_out/p2_ncu_super35.log 2>&1
ls -la gpurun_out
