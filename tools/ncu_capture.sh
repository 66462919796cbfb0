#!/bin/bash
# tools/ncu_capture.sh <tag> <kernel-regex>... : one `ncu --set full` capture per kernel of a one-iteration C4 optimize
# (tools/gpu_one.py c4 1), exported on the box as raw CSV (gpurun_out/<tag>_<kernel>.csv) -- the .ncu-rep files themselves
# (12 MB each) would not fit gpurun's 64 MiB return limit.  Run under gpurun after the plain command has exited 0.
tag=$1; shift
python tools/gpu_one.py c4 1 > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
for k in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o /tmp/${tag}_$k -f python tools/gpu_one.py c4 1 > gpurun_out/${tag}_ncu_$k.log 2>&1
  ncu -i /tmp/${tag}_$k.ncu-rep --page raw --csv > gpurun_out/${tag}_$k.csv 2>/dev/null
  echo "$k: $(wc -c < gpurun_out/${tag}_$k.csv) bytes"
  rm -f /tmp/${tag}_$k.ncu-rep
done
