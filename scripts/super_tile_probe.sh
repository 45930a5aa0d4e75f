# usage: super_tile_probe.sh tag...   (build/libqgmap_<tag>.so; "cur" = in-tree)
for t in "$@"; do
  if [ $t = cur ]; then python scripts/super_lanes_probe.py 2>&1 | grep "lanes=4" | sed "s/^/$t /";
  else QGMAP_LIB_PATH=build/libqgmap_$t.so python scripts/super_lanes_probe.py 2>&1 | grep "lanes=4" | sed "s/^/$t /"; fi
done
