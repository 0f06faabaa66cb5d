for ex in p2p allgather; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 4 --warmup 3 --exchange $ex > gpurun_out/ex_n2_$ex.json 2> gpurun_out/ex_n2_$ex.err
done
