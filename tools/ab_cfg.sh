#!/bin/bash
# tools/ab_cfg.sh <config> <variant>...  (on the GPU box): per-kernel times of bench config <config> for the in-tree lib and each variant
CFG=$1; shift
run() { python bench.py --config $CFG --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$CFG', '$1', round(d['ms_per_step'],3), {k: round(v*1e3) for k,v in d['roofline']['kernel_ms'].items()})"; }
run default
for v in "$@"; do HFR_B200_LIB=hifihr_b200/_build/lib_$v.so run $v; done
