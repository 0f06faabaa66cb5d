timeout 600 python tools/optix_compare.py > gpurun_out/optix_compare5.log 2>&1
timeout 300 python bench.py --steps 4 --warmup 3 > gpurun_out/bench_r01_synth.json 2> gpurun_out/bench_r01_synth.err
timeout 300 python bench.py --workload cornell --steps 4 --warmup 3 > gpurun_out/bench_r01_cornell.json 2> gpurun_out/bench_r01_cornell.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_ref.json 2> gpurun_out/bench_r01_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:pt_trace -c 64 --csv --log-file gpurun_out/trace_dram_bench_sg8.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_dram.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_bench_v5.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_v5.log 2>&1
