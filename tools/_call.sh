mkdir -p gpurun_out
timeout 120 xpic_b200/_build/fp64_peak > gpurun_out/r02_fp64_peak.json 2> gpurun_out/r02_fp64_peak.err
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "deposit or parity or golden or reproducible" > gpurun_out/r02_tests2.log 2>&1
rc=$?; echo "pytest rc $rc" >> gpurun_out/r02_tests2.log; tail -4 gpurun_out/r02_tests2.log
if [ $rc -ne 0 ]; then exit 1; fi
export XPIC_DEPOSIT_VARIANTS=0,3
timeout 600 python tools/profile_deposit.py > gpurun_out/r02_plain.log 2>&1 && \
XPIC_DEPOSIT_VARIANTS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cell_moments -s 2 -c 1 -f -o gpurun_out/r02_cell_moments_v2 python tools/profile_deposit.py > gpurun_out/r02_ncu.log 2>&1
cat gpurun_out/r02_fp64_peak.json gpurun_out/r02_plain.log; tail -3 gpurun_out/r02_ncu.log
