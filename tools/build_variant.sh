#!/bin/bash
# usage: tools/build_variant.sh NAME "-DFLAG=.." [header-with-plan-overrides]  -> birda_b200/variants/libbirda_b200_NAME.so
# The optional header is pre-included into k2_warp.cu (plan overrides hold commas, which -D cannot carry).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
D=$ROOT/birda_b200/variants/obj_$1
mkdir -p $D
cd $ROOT/birda_b200/csrc
INC=""
if [ -n "$3" ]; then INC="-include $3"; fi
for f in capi k1_pack k2_resample k2_warp k3_post k4_dense k5_melspec standin k6_flac; do
  if [ "$f" = "k2_warp" ] || [ ! -f $D/$f.o ]; then
    X=""; if [ "$f" = "k2_warp" ]; then X="$INC"; fi
    nvcc $2 $X -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -diag-suppress 20011,20014 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O2,-fno-fast-math -Xptxas -v -c $f.cu -o $D/$f.o 2> $D/$f.log &
  fi
done
g++ -O2 -std=c++17 -fPIC -fno-fast-math -ffp-contract=off -c rules.cpp -o $D/rules.o
g++ -O2 -std=c++17 -fPIC -c wav.cpp -o $D/wav.o
g++ -O2 -std=c++17 -fPIC -fno-fast-math -pthread -c pipeline.cpp -o $D/pipeline.o
g++ -O2 -std=c++17 -fPIC -pthread -c pool.cpp -o $D/pool.o
g++ -O2 -std=c++17 -fPIC -pthread -c watchdog.cpp -o $D/watchdog.o
g++ -O2 -std=c++17 -fPIC -c mask.cpp -o $D/mask.o
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $ROOT/birda_b200/variants/libbirda_b200_$1.so $D/*.o -cudart static
grep "Used" $D/k2_warp.log | tail -1
