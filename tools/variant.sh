#!/bin/bash
# usage: tools/variant.sh <name> <L> [-Dmacros...]   -- builds jwave-pro_b200/libjwavecuda_<name>.so: the current objects with
# jwc_dwt_fast.cu and jwc_modwt_fast.cu recompiled for ONE filter length and the given experiment macros
name=$1; L=$2; shift 2
d=jwave-pro_b200
objs=""
for f in $d/build/*.o; do case $f in *jwc_dwt_fast.o|*jwc_modwt_fast.o) ;; *) objs="$objs $f";; esac; done
for f in jwc_dwt_fast jwc_modwt_fast; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden,-O2 \
    --expt-relaxed-constexpr -DJWC_ONLY_L=$L "$@" -c $d/csrc/$f.cu -o /tmp/var_${name}_$f.o 2>/dev/null || exit 1
  objs="$objs /tmp/var_${name}_$f.o"
done
/usr/local/cuda/bin/nvcc -shared -o $d/libjwavecuda_$name.so $objs -gencode arch=compute_100a,code=sm_100a -lcudart_static -lpthread -ldl -lrt && echo built $d/libjwavecuda_$name.so
