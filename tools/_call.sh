python bench.py --gpus 1 --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/chk.json 2> gpurun_out/chk.err; echo rc=$?
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/chk.json') if l.startswith('{')][0])
print(d['ms_per_step'], d['roofline_dominant']['traffic'], d['roofline_dominant']['frac'], d['roofline']['frac'])
P
tail -3 gpurun_out/chk.err
