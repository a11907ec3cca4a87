mkdir -p gpurun_out
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29551 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_ecsim_n8.json 2> gpurun_out/r02_bench_n8.err; tail -c 600 gpurun_out/r02_bench_ecsim_n8.json; tail -3 gpurun_out/r02_bench_n8.err
timeout 600 $TR --nproc-per-node 4 --master-port 29552 bench.py --gpus 4 --steps 10 --warmup 3 --no-extra > gpurun_out/r02_bench_ecsim_n4.json 2> gpurun_out/r02_bench_n4.err; tail -c 300 gpurun_out/r02_bench_ecsim_n4.json
timeout 900 $TR --nproc-per-node 8 --master-port 29553 tools/spmv_sweep.py 8 512 > gpurun_out/r02_spmv_sweep_512_n8.json 2> gpurun_out/r02_spmv_sweep_512_n8.err; tail -3 gpurun_out/r02_spmv_sweep_512_n8.err | cut -c1-600
timeout 600 $TR --nproc-per-node 8 --master-port 29554 tests/multi_gpu_check.py > gpurun_out/r02_multi_gpu_check_n8.log 2>&1; tail -6 gpurun_out/r02_multi_gpu_check_n8.log
