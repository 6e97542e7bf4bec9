# End-of-round evidence on ONE B200 (under gpurun): tests, smoke, every bench config, launch list.  usage: bash tools/final_n1.sh <tag>
set -x
T=${1:-r2u}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${T}_pytest.log; cat gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 600 gpurun_out/${T}_bench.json
for c in c1 c3 c3m c5 c4 rp; do python bench.py --config $c --no-cpu-baseline --steps 10 --windows 3 > gpurun_out/${T}_bench_$c.json 2> gpurun_out/${T}_bench_$c.err; python -c "
import json
d=json.loads(open('gpurun_out/${T}_bench_$c.json').read())
print('$c', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['step_frac'], d['roofline']['frac'])
"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 3 --warmup 3 --windows 1 --no-cpu-baseline > gpurun_out/${T}_launches.log 2>&1
HFR_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_c3_launches.csv python bench.py --config c3 --steps 1 --warmup 3 --windows 1 --no-cpu-baseline > gpurun_out/${T}_c3_launches.log 2>&1
ls -la gpurun_out/${T}_*
