T=r2s
python -m pytest tests -m gpu -x -q -k "nimble or texel or pca or c3 or shade_backward_only or shader" 2>&1 | tail -4
for v in "" pca6 pca5 pca4; do
  if [ -n "$v" ]; then export HFR_B200_LIB=hifihr_b200/_build/lib_$v.so; fi
  python bench.py --config c3 --no-cpu-baseline --steps 10 --windows 3 > gpurun_out/${T}_bench_c3_$v.json 2> gpurun_out/${T}_bench_c3_$v.err; tail -2 gpurun_out/${T}_bench_c3_$v.err
  python -c "
import json
d=json.loads(open('gpurun_out/${T}_bench_c3_$v.json').read())
k=d['roofline']['kernel_ms']
print('variant [$v]', round(d['ms_per_step'],4), round(d['value']), 'shade_fwd', round(k['shade_fwd'],4), 'shade_raster_bwd', round(k['shade_raster_bwd'],4))
"
done
