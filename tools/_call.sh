set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_final_bench_ecsim_n8.json 2> gpurun_out/r02_final_bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 tests/multi_gpu_check.py > gpurun_out/r02_final_multi_gpu_check_n8.log 2>&1
tail -6 gpurun_out/r02_final_multi_gpu_check_n8.log
python - <<'P'
import json
for f in ('gpurun_out/r02_final_bench_ecsim_n8.json',):
    d=json.loads([l for l in open(f) if l.startswith('{')][0])
    print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], {k:v['ms_per_step'] for k,v in (d.get('other_configs') or {}).items()})
P
