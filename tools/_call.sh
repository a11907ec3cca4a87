set -x
python -m pytest tests/test_gpu_eccapfim.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1j.json 2> gpurun_out/r02_bench_n1j.err; python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_n1j.json') if l.startswith('{')][0])
print(d['ms_per_step'], d['e2e']['ms_per_step'], {k:v['ms_per_step'] for k,v in d['other_configs'].items()})
for k in d['kernels']: print(k['name'][:50], round(k['ms'],2), round(k.get('frac',0),3))
P
