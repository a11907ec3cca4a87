mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short --timeout 900 -k "deposit or decomposition or parity" > gpurun_out/r02_tests_n2f.log 2>&1; tail -5 gpurun_out/r02_tests_n2f.log
XPIC_DEPOSIT_VARIANTS=0,3 timeout 300 python tools/profile_deposit.py > gpurun_out/r02_plain2.log 2>&1; cat gpurun_out/r02_plain2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 --no-extra > gpurun_out/r02_bench_n2e.json 2> gpurun_out/r02_bench_n2e.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_n2e.json",):
    d=json.load(open(f)); print(f, d["ms_per_step"], d["config"]["stage_ms"], d["e2e"]["ms_per_step"]); print([(k["name"][:28], round(k["ms"],2)) for k in d["kernels"]]); print(d["roofline_dominant"]["family_ms_sum_vs_stage_clock"])
PY
