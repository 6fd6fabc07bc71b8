# usage: scratch/build_variants.sh name "flags" [name "flags" ...]
cd /root/repo/drone_b200/csrc
while [ $# -gt 0 ]; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared $2 -o ../../scratch/libs/lib_$1.so api.cu 2>&1 | grep -v "warning\|^$\|declared but never\|Remark\|\^" &
  shift 2
done
wait
ls -la /root/repo/scratch/libs
