python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 10 --warmup 3 --no-extra > gpurun_out/r02_final_bench_ecsim_n4.json 2> gpurun_out/r02_final_bench_n4.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r02_final_bench_ecsim_n4.json') if l.startswith('{')][0])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])
P
