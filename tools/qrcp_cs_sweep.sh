#!/bin/bash
# QRCP cluster-size sweep (tuning aid): bash tools/qrcp_cs_sweep.sh
for case in "1000 1" "1622 1" "2138 1" "520 14" "3120 5" "3120 18" "3120 36"; do
  set -- $case
  for cs in default 2 4 8 16; do
    if [ $cs = default ]; then unset ISDF_QR_CS; else export ISDF_QR_CS=$cs; fi
    echo -n "cs=$cs  "; python tools/qrcp_bench.py $1 $2 2>&1 | head -1
  done
done
