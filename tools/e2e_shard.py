"""What bounds the end-to-end step of a SHARD when all ranks of one box upload at once: run under
torchrun; every rank (1) times the pinned H2D copy of its shard's inputs alone and then concurrently with all other
ranks, (2) times the sharded one-call host API (`hostapi.histogram_loss_sharded`) at the chunk size given by
PH_HOST_CHUNK.  Rank 0 prints max-over-ranks figures.

    torchrun --nproc-per-node 8 tools/e2e_shard.py
"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch, torch.distributed as dist
import bench
from palette_and_histo_gan_b200 import _comm, hostapi

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
lo, hi = bench.shard_bounds(bench.GLOBAL_BATCH, world, rank)
real_np, fake_np, real_u8 = (a[lo:hi].copy() for a in bench.make_hist_inputs(bench.GLOBAL_BATCH, 47, with_u8=True))
real_h, fake_h = torch.from_numpy(real_u8).pin_memory(), torch.from_numpy(fake_np).pin_memory()
d_real, d_fake = torch.empty_like(real_h, device=dev), torch.empty_like(fake_h, device=dev)
nbytes = real_h.numel() + fake_h.numel() * 4

def h2d(n=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        d_real.copy_(real_h, non_blocking=True); d_fake.copy_(fake_h, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n

def maxr(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t)

h2d(2)
alone = None
for r in range(world):  # one rank at a time
    dist.barrier()
    if r == rank: alone = h2d()
dist.barrier(); together = h2d(); dist.barrier()
alone_ms, together_ms = maxr(alone) * 1e3, maxr(together) * 1e3
comm = _comm.peer_comm(True, dev)
ctx = hostapi.HostContext(lr)
grad_d = torch.empty((hi - lo, 64, 64, 4), dtype=torch.float32, device=dev)
step = lambda: hostapi.histogram_loss_sharded(comm, real_h, fake_h, bench.GLOBAL_BATCH, 64, out_grad_device=grad_d, ctx=ctx)
for _ in range(3): step()
dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): loss, _ = step()
torch.cuda.synchronize(); e2e_ms = maxr((time.perf_counter() - t0) / 10) * 1e3
if rank == 0:
    print(f"world {world} shard {hi - lo} images, {nbytes / 1e6:.1f} MB per rank and step; PH_HOST_CHUNK={os.environ.get('PH_HOST_CHUNK', 'default')}")
    print(f"  H2D alone    {alone_ms:.3f} ms = {nbytes / alone_ms / 1e6:.1f} GB/s per rank")
    print(f"  H2D together {together_ms:.3f} ms = {nbytes / together_ms / 1e6:.1f} GB/s per rank, {world * nbytes / together_ms / 1e6:.0f} GB/s box")
    print(f"  e2e step     {e2e_ms:.3f} ms = {bench.GLOBAL_BATCH / e2e_ms * 1e3:.0f} pairs/s (loss {loss:.9f})", flush=True)
dist.destroy_process_group()
