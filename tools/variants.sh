#!/bin/bash
# build variants of libmpp_b200.so with -D switches on the GPU box and time the colony pass with each
cd "$(dirname "$0")/.."
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off -shared"
i=0
for v in "$@"; do
  i=$((i+1))
  nvcc $FLAGS $v -o /tmp/libv$i.so maaco_path_planing_b200/csrc/*.cu || exit 1
  echo "== variant $i: $v"
  MPP_SO=/tmp/libv$i.so ${VCMD:-python tools/tour_time.py}
done
