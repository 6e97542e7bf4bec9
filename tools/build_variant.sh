#!/bin/bash
# tools/build_variant.sh <name> <extra nvcc flags...>  -> hifihr_b200/_build/lib_<name>.so  (tuning builds, git-ignored)
set -e
cd "$(dirname "$0")/.."
mkdir -p hifihr_b200/_build
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --shared -Xcompiler -fPIC -Xcompiler -ffp-contract=off \
  -I include -I hifihr_b200/csrc "$@" hifihr_b200/csrc/*.cu -o hifihr_b200/_build/lib_$name.so
echo built $name
