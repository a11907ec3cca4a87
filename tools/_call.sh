mkdir -p gpurun_out
make -C xpic_b200/host -s > /dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short --timeout 900 > gpurun_out/r02_tests_full_n2.log 2>&1; tail -30 gpurun_out/r02_tests_full_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 --no-extra > gpurun_out/r02_bench_n2b.json 2> gpurun_out/r02_bench_n2b.err; tail -c 1500 gpurun_out/r02_bench_n2b.json
