XPIC_SCHEME=ecsimcorr ncu --set full --clock-control none --import-source on -k regex:"k_esirkepov_cells" -s 2 -c 1 -o gpurun_out/r02_esirkepov python tools/profile_step.py 2 > gpurun_out/ncu_esir.log 2>&1
tail -3 gpurun_out/ncu_esir.log
