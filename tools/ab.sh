#!/bin/bash
# tools/ab.sh <tag> <variant>...  (on the GPU box): per-kernel times of the in-tree lib and of each tuning variant
TAG=$1; shift
mkdir -p gpurun_out
{
python tools/kernel_times.py
for v in "$@"; do python tools/kernel_times.py HFR_B200_LIB=hifihr_b200/_build/lib_$v.so; done
} > gpurun_out/${TAG}_ab.txt 2>&1
cat gpurun_out/${TAG}_ab.txt
