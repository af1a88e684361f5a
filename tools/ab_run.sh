#!/bin/bash
# A/B of library builds on the GPU box: the shipped library first, then every variant named on the command line
# (built beforehand with tools/build_variant.py; they travel with the snapshot).  Prints the stage timers of the
# second of two builds of the bench workload for each.
#   python tools/build_variant.py bk16 -DISDF_GEMM_3M_BK=16
#   gpurun --timeout 300 -- 'bash tools/ab_run.sh bk16'
WORKLOAD=${WORKLOAD:-diamond-standin-k333}
echo "== shipped"; python tools/profile_build.py "$WORKLOAD" 2>&1 | tail -1
for v in "$@"; do
  lib=$PWD/fft-isdf-scratch_b200/lib/variants/libisdf_b200_$v.so
  echo "== $v"; ISDF_B200_LIB=$lib python tools/profile_build.py "$WORKLOAD" 2>&1 | tail -1
done
