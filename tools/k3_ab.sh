#!/bin/bash
# K3 A/B on the GPU box: tools/k3_ab.sh "k3base default" [threads]  (variants built by tools/build_variant.sh; default = the
# in-tree library; threads = comma list of BIRDA_K3_THREADS values, default auto)
for v in $1; do
  echo "== variant $v"
  if [ "$v" = "default" ]; then unset BIRDA_B200_LIB; else export BIRDA_B200_LIB=$PWD/birda_b200/variants/libbirda_b200_$v.so; fi
  timeout 200 python tools/prof_k3.py ${2:-auto}
done
