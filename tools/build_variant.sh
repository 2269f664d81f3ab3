#!/bin/bash
# Developer A/B helper: builds mpc_rs_b200/libmpc_b200_<name>.so with extra -D flags applied to the MPPI kernels of
# model NL (FP32), everything else linked from the regular objects.  Select it at run time with MPCB_LIB_PATH.
#   tools/build_variant.sh <name> -DMPCB_SOMETHING=1 ...
set -e
name=$1; shift
cd "$(dirname "$0")/../mpc_rs_b200/csrc"
make -j8 >/dev/null
mkdir -p /tmp/mpcb_var_$name
for f in mppi_f32_NL mppi_f32x2_NL; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -ccbin g++ --cudart shared -Xcompiler -fPIC "$@" -c $f.cu -o /tmp/mpcb_var_$name/$f.o &
done
wait
objs=$(ls *.o | grep -v '^mppi_f32_NL.o$' | grep -v '^mppi_f32x2_NL.o$')
nvcc -gencode arch=compute_100a,code=sm_100a --cudart shared -shared -o ../libmpc_b200_$name.so $objs /tmp/mpcb_var_$name/mppi_f32_NL.o /tmp/mpcb_var_$name/mppi_f32x2_NL.o -ldl
echo built ../libmpc_b200_$name.so
