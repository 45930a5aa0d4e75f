# usage: ab_run.sh tag1 tag2 ...  (build/libqgmap_<tag>.so; "cur" = the in-tree library)
for t in "$@"; do
  if [ $t = cur ]; then python scripts/ab_bench.py cur short > gpurun_out/ab_$t.log 2>&1;
  else QGMAP_LIB_PATH=build/libqgmap_$t.so python scripts/ab_bench.py $t short > gpurun_out/ab_$t.log 2>&1; fi
done
for i in 1 2 3 4 5; do for t in "$@"; do sed -n ${i}p gpurun_out/ab_$t.log; done; done
