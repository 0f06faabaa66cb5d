for g in 8 16; do
for n in 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 4 --warmup 3 --sample-groups $g > gpurun_out/scale_n${n}_g${g}.json 2> gpurun_out/scale_n${n}_g${g}.err
done
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/scale_n2_g8.json 2> gpurun_out/scale_n2_g8.err
