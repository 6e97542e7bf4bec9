set -x
T=${1:-r2p}
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/${T}_pytest.log; cat gpurun_out/${T}_pytest.log
for c in c3 c3m; do python bench.py --config $c --no-cpu-baseline --steps 10 --windows 3 > gpurun_out/${T}_bench_$c.json 2> gpurun_out/${T}_bench_$c.err; tail -3 gpurun_out/${T}_bench_$c.err; python -c "
import json
d=json.loads(open('gpurun_out/${T}_bench_$c.json').read())
print('$c', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['step_frac'], d['roofline']['kernel_ms'], d['config'].get('launch'))
"; done
HFR_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_c3_launches.csv python bench.py --config c3 --steps 1 --warmup 3 --windows 1 --no-cpu-baseline > gpurun_out/${T}_c3_launches.log 2>&1
tail -2 gpurun_out/${T}_c3_launches.log | cut -c1-200
