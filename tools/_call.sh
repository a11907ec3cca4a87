set -x
XPIC_PROFILE_RANGE=1 python tools/profile_step.py 2 > gpurun_out/r02_plain5.log 2>&1 || exit 1
tail -1 gpurun_out/r02_plain5.log
XPIC_PROFILE_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_128x64.csv python tools/profile_step.py 2 > gpurun_out/r02_ncu5.log 2>&1
tail -2 gpurun_out/r02_ncu5.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
