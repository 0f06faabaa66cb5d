import os, pathlib, sys
import numpy as np, torch
import torch.distributed as dist
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from optix_raytracer_b200 import host
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = host.Context(local)
h, w = 64, 128
owner = torch.zeros((h, w, 4), dtype=torch.uint8, device=ctx.torch_device)
shared = host.share_result_buffer(ctx, owner, rank)
print(rank, "mapped", shared.device, hex(shared.data_ptr()), flush=True)
if rank == 1:
    shared[0].fill_(7)                       # executes in device 0's context of process 1
    torch.cuda.synchronize(0); print("fill via cuda:0 ok", flush=True)
    y = torch.full((w, 4), 9, dtype=torch.uint8, device="cuda:1")
    shared[1].copy_(y); torch.cuda.synchronize(); print("peer copy ok", flush=True)
    # a kernel of device 1 storing through the pointer: de-interleave of one fake gathered block into the shared frame
    g = torch.ones((1, w * h, 4), dtype=torch.float32, device="cuda:1")
    ctx.check(ctx.lib.b200rt_deinterleave(ctx.h, ctx.stream, g.data_ptr(), 1, w * h, w, h, 0, shared.data_ptr()), "deinterleave")
    torch.cuda.synchronize(); print("kernel store ok", flush=True)
dist.barrier()
if rank == 0:
    torch.cuda.synchronize()
    print("owner sees", owner[0, 0].tolist(), owner[1, 0].tolist(), owner[5, 5].tolist(), flush=True)
dist.barrier()
dist.destroy_process_group()
