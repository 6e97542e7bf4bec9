#!/bin/bash
# Run on the GPU box (under gpurun): launch list + full ncu capture of the dominant kernels.
# usage: tools/profile.sh <tag> [extra bench.py args, e.g. --config rp]
TAG=${1:-r2}
shift
EXTRA="$@"
OUT=gpurun_out
mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline $EXTRA > $OUT/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'raster_shade_fwd|raster_shade_pool|shade_bwd|loss_|mano_|geom_|skin_|blend_' -s 40 -c 16 \
    -o $OUT/${TAG}_prof -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline $EXTRA > $OUT/${TAG}_prof.log 2>&1
ls -la $OUT
