B200RT_PT_HOST_LOOP=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:pt_trace --csv --log-file gpurun_out/trace_dram_r02.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/trace_dram_r02.log 2>&1
grep -c pt_trace gpurun_out/trace_dram_r02.csv
