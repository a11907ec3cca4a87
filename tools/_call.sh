timeout 900 python -m pytest tests -x -q -m gpu -k "solve or gmres or state or energy or eccapfim or open" 2>&1 | tail -4
XPIC_BENCH_PRECOND=8 python tools/profile_step.py 3 2>&1 | tail -1
XPIC_DEPOSIT_VARIANTS=0 python tools/profile_deposit.py 2>/dev/null | tail -1
