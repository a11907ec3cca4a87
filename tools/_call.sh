set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "decomposition or two_gpus" 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-extra > gpurun_out/r02_bench_n2h.json 2> gpurun_out/r02_bench_n2h.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_n2h.json') if l.startswith('{')][0])
print(d['ms_per_step'], d['e2e']['ms_per_step'])
for k in d['kernels']: print(k['name'][:50], round(k['ms'],2), round(k.get('frac',0),3))
print(d['roofline_dominant']['family_ms_sum_vs_stage_clock'])
print(d['config']['stage_ms'])
P
