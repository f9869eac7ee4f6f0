"""End-to-end step of cfgC on one GPU for several host-pipeline chunk schedules (PH_HOST_CHUNK_LIST is read once per
process, so each schedule runs in a subprocess)."""
import sys, os, subprocess, time
sys.path.insert(0, os.getcwd())
SCHEDULES = ["", "148,296", "74,148,296", "148,222,296", "148,296,592,1184", "148,296,592,1184,1184,544,148",
             "222,444,888,888,888,544,222", "296,592,592,592,592,592,592,248"]
if len(sys.argv) > 1 and sys.argv[1] != "run":
    SCHEDULES = sys.argv[1:]
    sys.argv = sys.argv[:1]
if len(sys.argv) == 1:
    for sch in SCHEDULES:
        env = dict(os.environ)
        if sch: env["PH_HOST_CHUNK_LIST"] = sch
        out = subprocess.run([sys.executable, __file__, "run"], env=env, capture_output=True, text=True)
        print(f"{sch or 'default (uniform 296)':45s} {out.stdout.strip()} {out.stderr.strip()[-200:] if out.returncode else ''}", flush=True)
else:
    import torch, numpy as np, bench
    from palette_and_histo_gan_b200 import hostapi
    dev = torch.device("cuda:0")
    real_np, fake_np, real_u8 = bench.make_hist_inputs(4096, 47, with_u8=True)
    real_h = torch.from_numpy(real_u8).pin_memory(); fake_h = torch.from_numpy(fake_np).pin_memory()
    grad_d = torch.empty((4096, 64, 64, 4), dtype=torch.float32, device=dev)
    ctx = hostapi.HostContext(0)
    def step():
        s = hostapi.histogram_loss_begin(real_h, fake_h, 64, ctx=ctx)
        return hostapi.histogram_loss_finish(s, 4096, None, out_grad_device=grad_d, ctx=ctx)
    for _ in range(3): loss, _ = step()
    torch.cuda.synchronize(); ts = []
    for _ in range(8):
        t0 = time.perf_counter(); step(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    dt = float(np.median(ts))
    print(f"e2e {dt * 1e3:.2f} ms -> {4096 / dt:.0f} pairs/s (loss {loss:.9f})")
