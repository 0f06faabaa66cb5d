ncu --set full --clock-control none --import-source on -k regex:"radix_tree" -c 1 -o gpurun_out/prof_radix_r02b -f python tools/build_once.py --reps 1 > gpurun_out/ncu_radix.log 2>&1
tail -1 gpurun_out/ncu_radix.log
