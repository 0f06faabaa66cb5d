timeout 600 python -m pytest tests -x -q -m gpu > /tmp/o.txt 2>&1; echo "pytest rc=$?"; tail -6 /tmp/o.txt | cut -c1-300
timeout 600 python tools/optix_compare.py --skip-synth > gpurun_out/optix_compare7.log 2>&1; echo "compare rc=$?"
timeout 300 python bench.py --workload cornell --steps 4 --warmup 3 > gpurun_out/bench_r01_cornell.json 2> gpurun_out/bench_r01_cornell.err
