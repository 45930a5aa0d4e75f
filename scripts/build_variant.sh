#!/bin/bash
# build_variant.sh <tag> <extra nvcc flags...>: build libqgmap with extra -D flags into build/libqgmap_<tag>.so (A/B timing aid)
set -e
tag=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/gqmap-opticalflow_b200/csrc
obj=$root/build/obj_$tag
mkdir -p $obj
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -ccbin /usr/bin/g++"
$NV "$@" -Xptxas -v -c $src/qgmap_api.cu -o $obj/qgmap_api.o 2> $obj/ptxas.log
for f in qgmap_map qgmap_band qgmap_p2p qgmap_host; do [ -f $src/$f.o ] || make -C $src $f.o >/dev/null; done
$NV -shared -cudart shared -Xlinker -rpath=/usr/local/cuda/lib64 -o $root/build/libqgmap_$tag.so $obj/qgmap_api.o $src/qgmap_map.o $src/qgmap_band.o $src/qgmap_p2p.o $src/qgmap_host.o -ldl -lpthread -lrt
echo built $tag
