"""torchrun: raw NCCL all_to_all_single timing at the Ulysses message sizes (diagnostic)."""
import os, torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
for mb in (25, 75, 150, 300):
    n = mb * 1024 * 1024 // 2 // world * world
    a = torch.randn(n, device="cuda").bfloat16(); b = torch.empty_like(a)
    for _ in range(5): dist.all_to_all_single(b, a)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(20): dist.all_to_all_single(b, a)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    sent = n * 2 * (world - 1) / world
    if rank == 0: print(f"world {world} buffer {mb} MB: {ms:.3f} ms  -> {sent/ms/1e6:.1f} GB/s sent per rank", flush=True)
# plain copy for comparison
a = torch.randn(75 * 1024 * 1024 // 2, device="cuda").bfloat16(); b = torch.empty_like(a)
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
for _ in range(20): b.copy_(a)
e1.record(); torch.cuda.synchronize()
if rank == 0: print(f"local copy 75 MB: {e0.elapsed_time(e1)/20:.3f} ms")
dist.destroy_process_group()
