set -x
python bench.py --scheme eccapfim --steps 3 --warmup 3 --no-extra > gpurun_out/r02_bench_eccapfim_n1.json 2> gpurun_out/r02_bench_eccapfim_n1.err; python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_eccapfim_n1.json') if l.startswith('{')][0])
print(d['ms_per_step'], d.get('e2e',{}).get('ms_per_step'))
print(d.get('roofline')); print(d.get('kernels')); print(d['config'])
P
XPIC_SCHEME=eccapfim XPIC_BENCH_GRID=192,192,24 XPIC_BENCH_PPC=32 XPIC_PROFILE_RANGE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'k_cap_push_tasks' -c 2 -f -o gpurun_out/r02_cap_push python tools/profile_step.py 1 > gpurun_out/r02_ncu4.log 2>&1
tail -3 gpurun_out/r02_ncu4.log
