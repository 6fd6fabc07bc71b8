#!/bin/bash
# usage: profiles/tools/build_variants.sh name "flags" [name "flags" ...]
# A/B builds of libb200drone.so with -D knobs -> build/variants/lib_<name>.so; run with B2D_LIBRARY=<that file>.
cd /root/repo/drone_b200/csrc
mkdir -p ../../build/variants
while [ $# -gt 0 ]; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -pthread -shared -ldl -lpthread $2 -o ../../build/variants/lib_$1.so api.cu host_copy.o 2>&1 | grep -v "warning\|^$\|declared but never\|Remark\|\^" &
  shift 2
done
wait
ls -la /root/repo/build/variants
