#!/bin/bash
# development: build libpinsage_b200 with extra -D flags into gpurun_variants/<name>.so   usage: tools/build_variant.sh name [walker_source] -DFOO=1 ...
name=$1; shift
wsrc=gcn-song-embeddings_b200/csrc/walker.cu
if [ -f "$1" ]; then wsrc=$1; shift; fi
out=gpurun_variants; mkdir -p $out/obj_$name
cp $wsrc gcn-song-embeddings_b200/csrc/_walker_variant_$name.cu
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c gcn-song-embeddings_b200/csrc/_walker_variant_$name.cu -o $out/obj_$name/walker.o
rm gcn-song-embeddings_b200/csrc/_walker_variant_$name.cu
for f in gcn-song-embeddings_b200/build/*.o; do
  b=$(basename $f .o)
  [ "$b" == "walker" ] || cp $f $out/obj_$name/$b.o
done
nvcc -shared -o $out/$name.so $out/obj_$name/*.o -gencode arch=compute_100a,code=sm_100a && rm -rf $out/obj_$name && echo built $out/$name.so
