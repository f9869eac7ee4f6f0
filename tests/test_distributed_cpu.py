"""world_size-2 gloo test of the only exchange on the path: the all-reduce of the Hellinger sum of
squares (histogram.py:88-89) and the whole-batch size used by forward and backward."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from palette_and_histo_gan_b200 import histogram as H
    from oracle import histogram_oracle as ho

    rng = np.random.default_rng(11)
    real = np.tanh(rng.standard_normal((4, 8, 8, 4))).astype(np.float32)
    fake = np.tanh(rng.standard_normal((4, 8, 8, 4))).astype(np.float32)
    lo, hi = rank * 2, rank * 2 + 2
    local = ho.hist_loss_and_grad_f64(real[lo:hi], fake[lo:hi], size=16)
    ssum = torch.tensor([local["ssum"]], dtype=torch.float64)
    gb = H._reduce_over_ranks(ssum, 2, True, None)  # the product's own reduction helper
    whole = ho.hist_loss_and_grad_f64(real, fake, size=16)
    sharded = ho.hist_loss_and_grad_f64(real[lo:hi], fake[lo:hi], size=16, global_batch=gb, global_ssum=float(ssum))
    ok = (gb == 4 and abs(float(ssum) - whole["ssum"]) < 1e-12 and abs(sharded["loss"] - whole["loss"]) < 1e-14
          and np.allclose(sharded["grad"], whole["grad"][lo:hi], rtol=1e-10, atol=1e-16))
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_sharded_hellinger_reduction_gloo_world2():
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
        assert dict(out) == {0: True, 1: True}


def test_bench_shard_bounds():
    sys.path.insert(0, ROOT)
    import bench

    for total in (4096, 4097, 7):
        for world in (1, 2, 4, 8):
            spans = [bench.shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
