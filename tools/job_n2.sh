timeout 600 python -m pytest tests/test_gpu_async.py -m gpu -q -rs 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err; tail -c 600 gpurun_out/bench_r02_n2.json
