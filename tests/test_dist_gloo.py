"""World-size-2 gloo test of the host-side plumbing of the gathered global InfoNCE:
_dist_ctx / _gather_embeddings / _gather_lse_and_sums (the two collectives of a step), with the per-rank kernel math
stood in by the oracle (the CUDA kernels themselves are covered by tests/test_gpu_losses.py on one GPU with
emulated ranks).  Runs on CPU."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from clip_finegrained_alignment_b200 import losses as L
        from oracle import losses_oracle as lo
        B, D, s = 5, 12, 3.0
        g = torch.Generator().manual_seed(123)
        a_full = torch.randn(world * B, D, generator=g, dtype=torch.float64)
        b_full = torch.randn(world * B, D, generator=g, dtype=torch.float64)
        a, b = a_full[rank * B:(rank + 1) * B].contiguous(), b_full[rank * B:(rank + 1) * B].contiguous()
        w, r, grp = L._dist_ctx(None, True)
        assert (w, r) == (world, rank)
        assert L._dist_ctx(None, False) == (1, 0, None)
        a_all, b_all = L._gather_embeddings(torch.stack([a, b]), w, grp, False)
        assert torch.equal(a_all, a_full) and torch.equal(b_all, b_full)
        # raw layout handed to the tensor-core kernels: the untouched collective output [world, 2, B, D]
        ra, rb = L._gather_embeddings(torch.stack([a, b]), w, grp, True)
        raw4 = ra._base.view(world, 2, B, D)
        assert ra.data_ptr() == raw4.data_ptr() and rb.data_ptr() == raw4[0, 1].data_ptr()
        assert all(torch.equal(raw4[k, 0], a_full[k * B:(k + 1) * B]) and torch.equal(raw4[k, 1], b_full[k * B:(k + 1) * B])
                   for k in range(world))
        # per-rank "kernel" (oracle stand-in): local rows vs global columns, both directions
        fa, fb, bwd = lo.gathered_infonce_rank(a, b, a_all, b_all, rank, s, 0.5, 0.5)
        pack = torch.cat([fa["lse"], fb["lse"], torch.stack([fa["loss_sum"], fb["loss_sum"]])])   # kernel's [lse | sums]
        lse_all, sums = L._gather_lse_and_sums(pack, B, w, grp, False)   # [2, world*B] rank-major rows, global CE sums
        rawp, rawp2 = L._gather_lse_and_sums(pack, B, w, grp, True)        # raw [world, 2B+2] packs (kernel-side indexing)
        assert rawp is rawp2
        rp = rawp.view(world, 2 * B + 2)
        assert torch.equal(rp[:, :B].reshape(-1), lse_all[0]) and torch.equal(rp[:, B:2 * B].reshape(-1), lse_all[1])
        assert torch.allclose(rp[:, 2 * B:].sum(0), sums)
        loss = 0.5 * sums.sum() / (world * B)
        da, db = bwd(lse_all[0], lse_all[1])
        # single-process reference on the concatenated batch
        f1 = lo.infonce_forward(a_full, b_full, s)
        f2 = lo.infonce_forward(b_full, a_full, s)
        ref_loss = 0.5 * (f1["loss_sum"] + f2["loss_sum"]) / (world * B)
        da_ref, db_ref = lo.symmetric_infonce_backward(f1["ah"], f1["an"], f1["bh"], f1["bn"], f1["lse"], f2["lse"], s,
                                                       0.5, 0.5, float(world * B))
        ok = (abs(float(loss - ref_loss)) < 1e-12
              and torch.allclose(lse_all[0], f1["lse"]) and torch.allclose(lse_all[1], f2["lse"])
              and torch.allclose(da, da_ref[rank * B:(rank + 1) * B], atol=1e-12)
              and torch.allclose(db, db_ref[rank * B:(rank + 1) * B], atol=1e-12))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_gathered_infonce_plumbing_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, True), (1, True)]
