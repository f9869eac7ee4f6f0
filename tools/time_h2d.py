"""PCIe floor of the e2e step: pinned H2D of one step's inputs (uint8 real + float32 fake), and the e2e step at
several host-pipeline chunk sizes (PH_HOST_CHUNK is read once per process, so each size runs in a subprocess)."""
import sys, os, subprocess, time
sys.path.insert(0, os.getcwd())
if len(sys.argv) == 1:
    import torch
    dev = torch.device("cuda:0")
    a = torch.empty(4096 * 64 * 64 * 4, dtype=torch.float32).pin_memory(); b = torch.empty(4096 * 64 * 64 * 4, dtype=torch.uint8).pin_memory()
    da = torch.empty_like(a, device=dev); db = torch.empty_like(b, device=dev)
    for _ in range(2): da.copy_(a, non_blocking=True); db.copy_(b, non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): da.copy_(a, non_blocking=True); db.copy_(b, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"H2D {(a.nbytes + b.nbytes) / 1e6:.0f} MB: {dt * 1e3:.2f} ms = {(a.nbytes + b.nbytes) / dt / 1e9:.1f} GB/s", flush=True)
    for c in (148, 296, 444, 592, 1184):
        env = dict(os.environ, PH_HOST_CHUNK=str(c))
        print(c, subprocess.run([sys.executable, __file__, "run"], env=env, capture_output=True, text=True).stdout.strip(), flush=True)
else:
    import torch, numpy as np, bench
    from palette_and_histo_gan_b200 import hostapi
    dev = torch.device("cuda:0")
    real_np, fake_np = bench.make_hist_inputs(4096, 47)
    real_u8 = np.clip(np.rint((real_np + 1.0) * 127.5), 0, 255).astype(np.uint8)
    real_h = torch.from_numpy(real_u8).pin_memory(); fake_h = torch.from_numpy(fake_np).pin_memory()
    grad_d = torch.empty((4096, 64, 64, 4), dtype=torch.float32, device=dev)
    ctx = hostapi.HostContext(0)
    def step():
        s = hostapi.histogram_loss_begin(real_h, fake_h, 64, ctx=ctx)
        return hostapi.histogram_loss_finish(s, 4096, None, out_grad_device=grad_d, ctx=ctx)
    for _ in range(3): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(6): step()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 6
    print(f"e2e {dt * 1e3:.2f} ms -> {4096 / dt:.0f} pairs/s")
