#!/usr/bin/env python3
"""torchrun --nproc-per-node N tools/check_p2p_frame.py: the shared result buffer (host.share_result_buffer: every rank's launch stores its
pixels straight into rank 0's frame over NVLink) must equal the frame assembled from the all-gathered sample buffers, byte for byte."""
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
w, h = 640, 360
pt = host.PathTracer(ctx, w, h, 4, multigpu=(rank, world))
torch.cuda.synchronize(); print(rank, "built", flush=True)
pt.params.device_idx = 3          # no per-device tint (optixMultiGPU.cu:199-206), so the two assemblies are comparable
pt.sample_groups = 2
# (a) gather: compact per-rank sample buffers -> all-gather -> de-interleave (what bench.py does by default)
pt.launch_subframe(0)
torch.cuda.synchronize(); print(rank, "launched", flush=True)
gathered = torch.empty((world, pt.num_samples, 4), dtype=torch.float32, device=ctx.torch_device)
dist.all_gather_into_tensor(gathered.view(-1), pt.accum.view(-1))
torch.cuda.synchronize(); print(rank, "gathered", flush=True)
full_frame = torch.zeros((h, w, 4), dtype=torch.uint8, device=ctx.torch_device)
ctx.check(ctx.lib.b200rt_deinterleave(ctx.h, ctx.stream, gathered.data_ptr(), world, pt.num_samples, w, h, 0, full_frame.data_ptr()), "deinterleave")
torch.cuda.synchronize()
# (b) shared result buffer in rank 0's memory
shared = host.SharedResultBuffer(ctx, h, w, rank)
pt.params.result_buffer = shared.ptr
pt.launch_subframe(0)
torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    a, b = full_frame.cpu().numpy(), shared.tensor.cpu().numpy()
    assert a[..., :3].any(), "empty frame"
    assert np.array_equal(a, b), f"{(a != b).any(axis=-1).sum()} pixels differ between the gathered frame and the shared result buffer"
    print(f"P2P OK: {world} ranks, {w}x{h}, shared result buffer == gathered frame")
dist.barrier()
shared.close()
dist.destroy_process_group()
