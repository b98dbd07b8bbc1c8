#!/bin/bash
# build an experimental variant of the library: tools/build_variant.sh NAME -DDEVO_CORR_BOX=9 ...
name=$1; shift
out=devo_b200/lib/variants; mkdir -p $out/obj_$name
for f in devo_b200/csrc/*.cu; do
  b=$(basename $f .cu)
  if [ "$b" = "corr_fast" ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -Wno-deprecated-declarations "$@" -c $f -o $out/obj_$name/$b.o &
  else
    cp devo_b200/lib/obj/$b.o $out/obj_$name/$b.o
  fi
done
wait
nvcc -shared -o $out/libdevo_b200_$name.so $out/obj_$name/*.o -gencode arch=compute_100a,code=sm_100a -lcudart
echo built $out/libdevo_b200_$name.so
