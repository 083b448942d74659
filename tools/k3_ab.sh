#!/bin/bash
# K3 A/B on the GPU box: tools/k3_ab.sh "k3base k3p"  (variants built by tools/build_variant.sh)
for v in $1; do
  echo "== variant $v"
  export BIRDA_B200_LIB=$PWD/birda_b200/variants/libbirda_b200_$v.so
  timeout 120 python tools/prof_k3.py 2400 6522 sigmoid
  timeout 120 python tools/prof_k3.py 720 14795 softmax
  timeout 120 python tools/prof_k3.py 64 6522 sigmoid
done
