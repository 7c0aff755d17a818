#!/bin/bash
# usage: build_variants.sh name1:"-DX=1 -DY=2" name2:"" ...   -> build/var/lib_<name>.so (parallel)
mkdir -p build/var
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  ( /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
      -Xcompiler -fPIC -ccbin /usr/bin/g++ $flags -shared -o build/var/lib_$name.so mcstate_b200/csrc/api.cu \
      2>&1 | grep -i "error" ; echo "built $name" ) &
done
wait
