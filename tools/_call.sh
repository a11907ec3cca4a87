mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r02_bench_ecsim_n2.json 2> gpurun_out/r02_bench_n2.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_ecsim_n2.json",):
    d=json.load(open(f)); print(f, d["ms_per_step"], d["config"]["stage_ms"], d["e2e"]["ms_per_step"]); print([(k["name"][:28], round(k["ms"],2)) for k in d["kernels"]]); print(d["roofline_dominant"]["family_ms_sum_vs_stage_clock"]); print({k:(v.get("ms_per_step"),v.get("error")) for k,v in d["other_configs"].items()})
PY
