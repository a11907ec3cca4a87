mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_eccapfim.py -m gpu -q --tb=short --timeout 900 -k "solve or parity or golden or eccapfim or conservation" > gpurun_out/r02_tests_cmd.log 2>&1; tail -6 gpurun_out/r02_tests_cmd.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r02_bench_n1d.json 2> gpurun_out/r02_bench_n1d.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_n1d.json")); print(d["ms_per_step"], d["config"]["stage_ms"]); print([(k["name"][:28], round(k["ms"],2), round(k.get("frac") or 0,3)) for k in d["kernels"]]); print(d["roofline_dominant"]["family_ms_sum_vs_stage_clock"])
PY
