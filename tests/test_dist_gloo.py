"""world_size-2 tests of the data-parallel statistics protocol on CPU (gloo).  The device kernels
are replaced by tiny torch stand-ins injected through the `_k` hook, so what is tested is the host
logic that ships: which collectives run, in which order, and that the sharded result equals the
single-process result on the concatenated batch."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import losshead_oracle as O


class CpuKernels:
    """torch stand-ins for the handful of ops.* kernels the protocol calls."""

    @staticmethod
    def cols_lse(x, inv_tau, rowbias=None):
        u = x.float() * inv_tau
        if rowbias is not None:
            u = u - rowbias[:, None]
        return torch.logsumexp(u, dim=0)

    @staticmethod
    def rows_lse(x, inv_tau, colbias=None):
        u = x.float() * inv_tau
        if colbias is not None:
            u = u - colbias[None, :]
        return torch.logsumexp(u, dim=1)

    @staticmethod
    def axpb(a, alpha, beta=0.0, out=None):
        return a * alpha + beta

    @staticmethod
    def lse_combine(gathered, add=0.0):
        return torch.logsumexp(gathered, dim=0) + add

    @staticmethod
    def cols_sum(x):
        return x.float().sum(0)

    @staticmethod
    def center_ema_(center, colsum, global_rows, momentum):
        center.copy_(center * momentum + (colsum / global_rows).reshape(1, -1) * (1 - momentum))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dinox_b200 import losshead
    g = torch.Generator().manual_seed(123)
    K, rows = 96, 8
    t_all = torch.randn(world * rows, K, generator=g) * 1.5
    t_loc = t_all[rank * rows:(rank + 1) * rows].contiguous()
    # --- Sinkhorn: sharded == global
    a, b = losshead.sinkhorn_knopp_biases(t_loc, 0.04, 3, process_group=True, _k=CpuKernels)
    q_loc = torch.exp(t_loc / 0.04 - a[None, :] - b[:, None])
    q_ref = O.sinkhorn_knopp(t_all, 0.04, 3)[rank * rows:(rank + 1) * rows]
    err_sk = ((q_loc - q_ref).norm() / q_ref.norm()).item()
    # --- centre: all-reduced column sums / global rows == single-process mean
    dl = losshead.DINOLoss(K, 0.9, process_group=True)
    dl.update_center(t_loc, _k=CpuKernels)
    c_ref = O.center_update(torch.zeros(1, K), t_all, 0.9)
    err_c = ((dl.center - c_ref).norm() / c_ref.norm()).item()
    # --- sum all-reduce helper
    v = torch.full((4,), float(rank + 1))
    losshead.allreduce_sum_(v, True)
    ok_sum = bool((v == sum(range(1, world + 1))).all())
    ret[rank] = (err_sk, err_c, ok_sum)
    dist.barrier()
    dist.destroy_process_group()


def test_dp_statistics_world2_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for r in range(world):
        err_sk, err_c, ok_sum = ret[r]
        assert err_sk < 1e-4, err_sk
        assert err_c < 1e-6, err_c
        assert ok_sum


def test_single_process_protocol_equals_oracle():
    from dinox_b200 import losshead
    g = torch.Generator().manual_seed(5)
    t = torch.randn(12, 64, generator=g)
    a, b = losshead.sinkhorn_knopp_biases(t, 0.04, 3, process_group=None, _k=CpuKernels)
    q = torch.exp(t / 0.04 - a[None, :] - b[:, None])
    assert torch.allclose(q, O.sinkhorn_knopp(t, 0.04, 3), rtol=1e-4, atol=1e-7)
    assert torch.allclose(q.sum(-1), torch.ones(12), rtol=1e-5)


def test_reference_arm_rank_gating(monkeypatch, capsys):
    """bench.py --impl reference: only rank 0 prints, other ranks exit without work."""
    import importlib
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    bench = importlib.import_module("bench")
    monkeypatch.setenv("RANK", "1")

    class A:
        config, cpu_sample_batch, steps, warmup, accum, gpus = "C2", 1, 1, 0, 4, 2
    bench.run_reference(A)
    assert capsys.readouterr().out.strip() == ""
