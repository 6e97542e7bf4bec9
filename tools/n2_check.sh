python -m pytest tests -m gpu -x -q -k "two_gpu" 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2u_bench_c2_n2.json 2> gpurun_out/r2u_bench_c2_n2.err; tail -2 gpurun_out/r2u_bench_c2_n2.err; python -c "
import json
d=json.loads(open('gpurun_out/r2u_bench_c2_n2.json').read()); print('c2 n2', d['ms_per_step'], d['value'], d['e2e']['value'], d['n_gpus'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config c3 --steps 10 --warmup 3 --windows 3 > gpurun_out/r2u_bench_c3_n2.json 2> gpurun_out/r2u_bench_c3_n2.err; tail -2 gpurun_out/r2u_bench_c3_n2.err; python -c "
import json
d=json.loads(open('gpurun_out/r2u_bench_c3_n2.json').read()); print('c3 n2', d['ms_per_step'], d['value'], d['e2e']['value'], d['n_gpus'])"
