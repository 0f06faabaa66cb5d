timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "cornell_pathtracer" > /tmp/o.txt 2>&1; echo "rc=$?"; tail -5 /tmp/o.txt | cut -c1-300
for rs in 0 1; do timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --ray-sort $rs > gpurun_out/rs_$rs.json 2> gpurun_out/rs_$rs.err; done
