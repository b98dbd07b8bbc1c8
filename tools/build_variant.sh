#!/bin/bash
# build an experimental variant of the library: tools/build_variant.sh NAME FILE.cu -DDEVO_CORR_BOX=9 ...
# (only FILE.cu is recompiled with the extra flags; everything else is linked from the main build)
name=$1; file=$2; shift 2
out=devo_b200/lib/variants; mkdir -p $out/obj_$name
cp devo_b200/lib/obj/*.o $out/obj_$name/
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -Wno-deprecated-declarations "$@" \
  -c devo_b200/csrc/$file -o $out/obj_$name/$(basename $file .cu).o || exit 1
nvcc -shared -o $out/libdevo_b200_$name.so $out/obj_$name/*.o -gencode arch=compute_100a,code=sm_100a -lcudart
rm -rf $out/obj_$name
echo built $out/libdevo_b200_$name.so
