set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "deposit or batched or staging" 2>&1 | tail -3
XPIC_DEPOSIT_VARIANTS=0,4 python tools/profile_deposit.py > gpurun_out/prof_tiles.json 2> gpurun_out/prof_tiles.err
cat gpurun_out/prof_tiles.json
XPIC_WS_PROF=1 XPIC_DEPOSIT_VARIANTS=0 python tools/profile_deposit.py > gpurun_out/prof_ws.json 2> gpurun_out/prof_ws.err
tail -3 gpurun_out/prof_ws.err
