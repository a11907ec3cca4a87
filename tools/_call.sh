mkdir -p gpurun_out
make -C xpic_b200/host -s > /dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short --timeout 900 > gpurun_out/r02_tests_full_n2c.log 2>&1; tail -12 gpurun_out/r02_tests_full_n2c.log
timeout 1200 python tools/spmv_sweep.py 8 64,128,256 > gpurun_out/r02_spmv_sweep_1gpu.json 2> gpurun_out/r02_spmv_sweep_1gpu.err; tail -4 gpurun_out/r02_spmv_sweep_1gpu.err
