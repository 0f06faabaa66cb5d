timeout 300 python tools/profile_step.py --triangles 50000000 --width 1920 --height 1080 --spl 4 --steps 2 --sample-groups 4 > gpurun_out/plain_prof.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pt_trace_kernel -s 1 -c 1 -f -o gpurun_out/prof_trace_r01f python tools/profile_step.py --triangles 50000000 --width 1920 --height 1080 --spl 4 --steps 1 --sample-groups 4 > gpurun_out/ncu_prof.log 2>&1
tail -2 gpurun_out/ncu_prof.log
