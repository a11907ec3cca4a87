mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short --timeout 120 -k "deposit or parity or golden or reproducible or decomposition" > gpurun_out/r02_tests3.log 2>&1
rc=$?; echo "pytest rc $rc" >> gpurun_out/r02_tests3.log; tail -15 gpurun_out/r02_tests3.log
if [ $rc -ne 0 ]; then exit 1; fi
export XPIC_DEPOSIT_VARIANTS=0,3
timeout 300 python tools/profile_deposit.py > gpurun_out/r02_plain.log 2>&1 && \
XPIC_DEPOSIT_VARIANTS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cell_moments -s 2 -c 1 -f -o gpurun_out/r02_cell_moments_ws python tools/profile_deposit.py > gpurun_out/r02_ncu.log 2>&1
cat gpurun_out/r02_plain.log; tail -3 gpurun_out/r02_ncu.log
