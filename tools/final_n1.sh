set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2l_pytest.log; cat gpurun_out/r2l_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; tail -c 600 gpurun_out/r2l_bench.json
for c in c1 c3 c5 c4 rp; do python bench.py --config $c --no-cpu-baseline --steps 10 --windows 3 > gpurun_out/r2l_bench_$c.json 2> gpurun_out/r2l_bench_$c.err; python -c "
import json
d=json.loads(open('gpurun_out/r2l_bench_$c.json').read())
print('$c', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['step_frac'], d['roofline']['frac'])
"; done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2l_bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/r2l_bench_reference.json
python bench.py --impl reference --config c1 --steps 3 --warmup 1 > gpurun_out/r2l_bench_reference_c1.json 2>/dev/null; cut -c1-200 gpurun_out/r2l_bench_reference_c1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2l_launches.csv python bench.py --steps 3 --warmup 3 --windows 1 --no-cpu-baseline > gpurun_out/r2l_launches.log 2>&1
HFR_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:'raster_shade_fwd|shade_bwd_tiled|loss_fwd|loss_bwd|geom_|hand_|blend_|skin_|raster_setup|raster_order|rec_gather' -s 40 -c 19 -o gpurun_out/r2l_prof -f python bench.py --steps 3 --warmup 3 --windows 1 --no-cpu-baseline > gpurun_out/r2l_prof.log 2>&1
ls -la gpurun_out/r2l_*
