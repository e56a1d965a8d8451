"""N > 1 host logic on CPU: world_size-2 gloo.  The ranks' partial reduced forms (computed here
by the oracle, as K3 would on each rank's slice of the source) are all-reduced and must equal the
unsharded reduction; pair ranges must tile the batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _partial_form(src, q, W, mu):
    """sum p~p~^T (x) W, sum (W e) p~^T, sum e^T W e at T = I (dense, float64)."""
    n, d = src.shape
    e = q - src
    pt = np.concatenate([np.ones((n, 1)), src - mu], axis=1)
    H = np.einsum("na,nb,ncd->abcd", pt, pt, W)
    G = np.einsum("ncd,nd,na->ca", W, e, pt)
    c = np.einsum("ni,nij,nj->", e, W, e)
    return np.concatenate([H.ravel(), G.ravel(), [c, float(n)]])


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, ROOT)
    from generalized_icp_b200 import sharding
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)
    n = 1001
    src = rng.normal(size=(n, 3)) * 5
    q = src + rng.normal(size=(n, 3)) * 0.1
    A = rng.normal(size=(n, 3, 3))
    W = A @ A.transpose(0, 2, 1) + np.eye(3)
    mu = src.mean(0)
    lo, hi = sharding.source_slice(n, rank, world)
    part = _partial_form(src[lo:hi], q[lo:hi], W[lo:hi], mu)
    total = sharding.allreduce_reduced_form(part).numpy()
    full = _partial_form(src, q, W, mu)
    ok = np.allclose(total, full, rtol=1e-12, atol=1e-9)
    # pair sharding tiles the batch
    ranges = [sharding.pair_range(4097, r, world) for r in range(world)]
    tiles = ranges[0][0] == 0 and ranges[-1][1] == 4097 and all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
    eq = [sharding.equal_slice(n, r, world) for r in range(world)]
    eq_ok = eq[0][0] == 0 and eq[-1][1] == n and all(e[1] - e[0] <= (n + world - 1) // world for e in eq)
    out[rank] = int(ok and tiles and eq_ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_allreduce_of_partial_forms(world):
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert [out[r] for r in range(world)] == [1] * world
