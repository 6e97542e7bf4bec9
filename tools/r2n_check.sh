set -x
T=${1:-r2q}
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/${T}_pytest.log; cat gpurun_out/${T}_pytest.log
for c in c3 rp c1; do python bench.py --config $c --no-cpu-baseline --steps 10 --windows 3 > gpurun_out/${T}_bench_$c.json 2> gpurun_out/${T}_bench_$c.err; tail -3 gpurun_out/${T}_bench_$c.err; python -c "
import json
d=json.loads(open('gpurun_out/${T}_bench_$c.json').read())
print('$c', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['step_frac'], d['roofline']['kernel_ms'], d['config'].get('launch'))
"; done
