"""Diagnostic (not a test): all_gather_into_tensor bandwidth for the layer-exchange size. Run under torchrun."""
import os, sys, time
import torch, torch.distributed as dist
world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
for dt, name in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
    n_max = (690599 + world - 1) // world
    loc = torch.randn(n_max, 128, device=dev).to(dt)
    out = torch.empty(world * n_max, 128, device=dev, dtype=dt)
    for _ in range(5): dist.all_gather_into_tensor(out, loc)
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): dist.all_gather_into_tensor(out, loc)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    recv = out.numel() * out.element_size() * (world - 1) / world
    if rank == 0: print(f"world={world} {name} all_gather {out.numel()*out.element_size()/1e6:.0f} MB: {ms:.3f} ms, recv {recv/ms/1e6:.0f} GB/s per rank, env={os.environ.get('PROBE_TAG','default')}", flush=True)
dist.barrier(); dist.destroy_process_group()
