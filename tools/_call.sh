# scratch command file for `gpurun -- 'bash tools/_call.sh'`; the last content: the checks of the final round-2 build
set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 400 gpurun_out/bench_n1.json
