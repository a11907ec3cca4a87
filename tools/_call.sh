set -x
python -m pytest tests/test_gpu_eccapfim.py -x -q -m gpu 2>&1 | tail -5
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "step_host_equals" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1f.json 2> gpurun_out/r02_bench_n1f.err; python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_n1f.json') if l.startswith('{')][0])
print(d['ms_per_step'], d['e2e'], {k:(v['ms_per_step']) for k,v in d['other_configs'].items()})
P
