timeout 900 python tools/optix_compare.py > gpurun_out/optix_compare8.log 2>&1; echo "compare rc=$?"
timeout 300 python bench.py --steps 4 --warmup 3 > gpurun_out/bench_r01_synth.json 2> gpurun_out/bench_r01_synth.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_ref.json 2> gpurun_out/bench_r01_ref.err
