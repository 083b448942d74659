#!/bin/bash
# A/B kernel variants on the GPU box: tools/ab.sh "u1 u2 t1" [minutes] [config]
for v in $1; do
  echo "== variant $v"
  BIRDA_B200_LIB=$PWD/birda_b200/variants/libbirda_b200_$v.so timeout 120 python tools/ab_check.py ${3:-c2}
  BIRDA_B200_LIB=$PWD/birda_b200/variants/libbirda_b200_$v.so timeout 120 python tools/prof_run.py ${2:-60} 5 ${3:-c2}
done
