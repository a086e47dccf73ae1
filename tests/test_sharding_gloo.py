"""world_size-2 (gloo, CPU) check of the multi-GPU decomposition used by engine.PGDEngine: every rank owns a
contiguous range of tile rows, computes the direct AND mirrored contribution of its tiles into a full-length
partial result, and the partials are summed with all_reduce; the degree vector uses the d_fill protocol
(1 on rank 0, 0 elsewhere).  The per-shard arithmetic here is numpy (the CUDA kernels need a GPU); what is under test
is the host-side sharding / reduction logic that engine.py drives."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

TILE = 128


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, K, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from mcgra_b200.engine import shard_tile_rows
    rng = np.random.RandomState(0)
    A = np.tril(rng.random_sample((n, n)).astype(np.float32), -1)      # strict lower triangle (the stored part)
    B = rng.standard_normal((n, K)).astype(np.float32)
    T = (n + TILE - 1) // TILE
    tr0, tr1 = shard_tile_rows(T, world)[rank]
    Y = np.zeros((n, K), np.float64)
    d = np.full(n, 1.0 if rank == 0 else 0.0)
    for I in range(tr0, tr1):
        for J in range(I + 1):
            i0, i1, j0, j1 = I * TILE, min(n, (I + 1) * TILE), J * TILE, min(n, (J + 1) * TILE)
            Tt = A[i0:i1, j0:j1].astype(np.float64)
            Y[i0:i1] += Tt @ B[j0:j1]            # direct product
            Y[j0:j1] += Tt.T @ B[i0:i1]          # mirrored product
            d[i0:i1] += Tt.sum(1)
            d[j0:j1] += Tt.sum(0)
    Yt, dt = torch.from_numpy(Y), torch.from_numpy(d)
    dist.all_reduce(Yt)
    dist.all_reduce(dt)
    if rank == 0:
        S = (A + A.T).astype(np.float64)
        out["y_err"] = float(np.max(np.abs(Yt.numpy() - S @ B)))
        out["d_err"] = float(np.max(np.abs(dt.numpy() - (1.0 + S.sum(1)))))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [300, 777])
def test_tile_row_sharding_reduces_to_full_product(n):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, 8, out), nprocs=world, join=True)
    assert out["y_err"] < 1e-9 and out["d_err"] < 1e-9
