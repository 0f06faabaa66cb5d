set -x
for g in 4 8 16; do
  timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sample-groups $g > gpurun_out/sg_synth_$g.json 2> gpurun_out/sg_synth_$g.err
done
for g in 1 4 8 16; do
  timeout 300 python bench.py --workload cornell --steps 4 --warmup 3 --no-cpu-baseline --sample-groups $g > gpurun_out/sg_cornell_$g.json 2> gpurun_out/sg_cornell_$g.err
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:pt_trace -c 48 --csv --log-file gpurun_out/trace_dram_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_dram.log 2>&1
