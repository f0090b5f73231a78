"""CPU, world_size 2 over gloo: the multi-GPU exact mode's host-side logic.  Each rank takes the
tile range cge_b200_shard_plan gives it, forms the partial degree sums of exactly those tiles
(numpy stands in for the sweep kernel), and one all-reduce per pass must reproduce the
single-rank sums -- the exchange step of SURVEY.md 8(e)."""
import os
import socket

import numpy as np
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


def _tile_list(nb):
    return [(bi, bj) for bi in range(nb) for bj in range(bi, nb)]


def _partial_sums(G, T, tiles):
    """What the sweep kernel leaves in part[][] for `tiles`, already summed over blocks."""
    n = T.shape[0]
    s = np.zeros(n)
    for bi, bj in tiles:
        r = slice(bi * TILE, min((bi + 1) * TILE, n))
        c = slice(bj * TILE, min((bj + 1) * TILE, n))
        blk = G[r, c]
        s[r] += blk @ T[c]
        if bi != bj:
            s[c] += blk.T @ T[r]
    return s


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cge_jl_b200 import _lib

    rng = np.random.default_rng(0)       # same problem on every rank
    x = rng.normal(size=(n, 5))
    D = np.sqrt(((x[:, None] - x[None]) ** 2).sum(-1))
    G = (1 - D / D.max()) ** 1.5
    w = rng.uniform(1, 5, n)
    T = np.ones(n)
    nb = (n + TILE - 1) // TILE
    nt, b, e = _lib.shard_plan(n, rank, world)
    mine = _tile_list(nb)[b:e]
    diffs = []
    for _ in range(5):                   # five passes of divergence.jl:152-166
        sraw = torch.from_numpy(_partial_sums(G, T, mine))
        dist.all_reduce(sraw)            # the per-pass exchange
        S = T * sraw.numpy()
        T = T + 0.25 * T * (w / S - 1.0)
        diffs.append(np.abs(w - S).max())
    if rank == 0:
        q.put((nt, diffs, T))
    dist.destroy_process_group()


def test_two_rank_pass_matches_single_rank():
    n, world = 300, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    nt, diffs, T = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-rank reference of the same five passes
    rng = np.random.default_rng(0)
    x = rng.normal(size=(n, 5))
    D = np.sqrt(((x[:, None] - x[None]) ** 2).sum(-1))
    G = (1 - D / D.max()) ** 1.5
    w = rng.uniform(1, 5, n)
    T1 = np.ones(n)
    d1 = []
    for _ in range(5):
        S = T1 * (G @ T1)
        T1 = T1 + 0.25 * T1 * (w / S - 1.0)
        d1.append(np.abs(w - S).max())
    assert nt == 6
    np.testing.assert_allclose(diffs, d1, rtol=1e-12)
    np.testing.assert_allclose(T, T1, rtol=1e-12)
