set -x
python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r02_final_bench_n1.json 2> gpurun_out/r02_final_bench_n1.err
tail -c 600 gpurun_out/r02_final_bench_n1.json
XPIC_PROFILE_RANGE=1 XPIC_BENCH_PRECOND=8 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_final_launches.csv python tools/profile_step.py 2 > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
XPIC_DEPOSIT_VARIANTS=0 ncu --set full --clock-control none --import-source on -k regex:"k_cell_moments_ws|k_gather_tiles" -s 30 -c 3 -o gpurun_out/r02_final_moments python tools/profile_deposit.py > gpurun_out/ncu_tiles.log 2>&1
tail -3 gpurun_out/ncu_tiles.log
