"""GPU (>= 2 devices): the sharded exact mode.  2, 4 or 8 ranks (one process per GPU, as many as the
box has) own contiguous shares of the tile sequence (stored regime) or of the super-tile sequence
(recompute regime) and exchange the partial degree sums every pass -- over NCCL from the host
loop, or over NVLink peer memory inside the persistent kernel; the result must match the
single-GPU run (identical pass counts and best alphas, scores within 1e-9) and the CPU oracle.
Cases that need more GPUs than the box has are skipped; run with `gpurun --gpus N`."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem(directed):
    from util import planted_partition
    n = 1500
    edges, ew, vw, comm, emb = planted_partition(n, 9, 24, seed=77, directed=directed, weighted=True)
    return n, edges, ew, vw, comm, emb


def _worker(rank, world, port, directed, p2p, regime, q, fuse=False):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if fuse:  # per-alpha kernels and the deferred B sweep (k_bfp), as on large problems
        os.environ.update(CGE_B200_RT_EXPONENT="0", CGE_B200_FUSE_B="1")
    if regime == 2:
        os.environ["CGE_B200_RC_SB"] = "2"  # super-tiles of 2 x 2 tiles: 21 work units over the ranks
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    from cge_jl_b200 import divergence as dv
    n, edges, ew, vw, comm, emb = _problem(directed)
    samples = dv.draw_samples(edges, ew, n, 2000, 42, directed, True)
    sc = dv.Scorer(rank)
    ids = [dv.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    sc.comm_init(ids[0], rank, world)
    if p2p:  # NVLink peer exchange inside the persistent kernel instead of NCCL per pass
        handles = [None] * world
        dist.all_gather_object(handles, sc.p2p_export(n))
        sc.p2p_import(handles)
    p, keep = dv.make_problem(edges, ew, comm, emb, np.zeros(n), vw, None, None, None, False,
                              directed, samples, 0, 0, regime)
    sc.upload(p, keep)
    out, st = sc.run()
    dist.barrier()
    if rank == 0:
        q.put((out, list(st.iters), list(st.div), list(st.auc), int(st.n_ranks), int(st.driver),
               int(st.regime), int(st.b_fused)))
    sc.close()
    dist.destroy_process_group()


CASES = [(2, False, False, 1), (2, True, False, 1), (2, False, True, 1), (2, True, True, 1),
         (2, False, True, 2), (2, True, True, 2), (2, False, False, 2),
         (4, False, True, 1), (4, True, True, 2), (8, False, True, 1), (8, False, True, 2),
         (8, True, True, 1)]


@pytest.mark.parametrize("world,directed,p2p,regime", CASES,
                         ids=[f"{w}gpu-{'dir' if d else 'undir'}-{'nvlink' if p else 'nccl'}-"
                              f"{'stored' if r == 1 else 'recompute'}" for w, d, p, r in CASES])
def test_sharded_matches_one_gpu_and_oracle(world, directed, p2p, regime):
    """The tile (stored) or super-tile (recompute) sequence sharded over `world` ranks, per-pass
    exchange through NCCL from the host loop or over NVLink peer memory inside the persistent
    kernel: identical pass counts and best alphas, scores within 1e-9 of the single-GPU run and of
    the oracle."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp

    import oracle
    from cge_jl_b200 import divergence as dv
    from util import RTOL, assert_parity, empty_landmark_args

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, directed, p2p, regime, q))
             for r in range(world)]
    for p in procs:
        p.start()
    out2, iters2, div2, auc2, n_ranks, driver, reg, _ = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert n_ranks == world and driver == (2 if p2p else 1) and reg == regime
    n, edges, ew, vw, comm, emb = _problem(directed)
    samples = dv.draw_samples(edges, ew, n, 2000, 42, directed, True)
    f = dv.wGCL_directed if directed else dv.wGCL
    out1, st1 = f(edges, ew, comm, emb, np.zeros(n), vw, *empty_landmark_args(), False, 42, 2000,
                  False, samples=samples, return_stats=True)
    assert list(st1.iters) == iters2
    assert out1[0] == out2[0] and out1[4] == out2[4]
    np.testing.assert_allclose(out2, out1, rtol=RTOL, atol=1e-15)
    fo = oracle.wgcl_directed if directed else oracle.wgcl
    ref, tr = fo(edges, ew, comm, emb, np.zeros(n), vw, samples=samples)

    class St:
        n_alpha_run = tr.n_alpha_run
        iters, div, auc = iters2, div2, auc2
    assert_parity(out2, St, ref, tr)


@pytest.mark.parametrize("regime,driver", [(1, 0), (2, 0), (1, 1)],
                         ids=["stored", "recompute", "stored-hostloop-asked"])
@pytest.mark.parametrize("directed", [False, True])
def test_single_process_two_gpus(directed, regime, driver):
    """cge_b200_score_multi: the ranks are threads of one process (the Julia ccall case); peers
    are reached by plain peer access, extrema and B are reduced on the host.  There is no NCCL
    communicator on this path, so a request for the host-loop driver still runs the persistent
    kernel (it used to dereference the missing communicator)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from cge_jl_b200 import divergence as dv
    from util import RTOL, empty_landmark_args

    n, edges, ew, vw, comm, emb = _problem(directed)
    samples = dv.draw_samples(edges, ew, n, 2000, 42, directed, True)
    p, keep = dv.make_problem(edges, ew, comm, emb, np.zeros(n), vw, None, None, None, False,
                              directed, samples, 0, driver, regime)
    out2, st2 = dv.score_multi(p, 2)
    assert st2.n_ranks == 2 and st2.driver == 2 and st2.regime == regime
    f = dv.wGCL_directed if directed else dv.wGCL
    out1, st1 = f(edges, ew, comm, emb, np.zeros(n), vw, *empty_landmark_args(), False, 42, 2000,
                  False, samples=samples, return_stats=True)
    assert list(st1.iters) == list(st2.iters)
    assert out1[0] == out2[0] and out1[4] == out2[4]
    np.testing.assert_allclose(out2, out1, rtol=RTOL, atol=1e-15)
    del keep


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_deferred_b_sweep(world, monkeypatch):
    """The B sweep of an alpha riding on the first pass of the next one (k_bfp) with the tiles sharded over
    `world` ranks: every rank fills its own partial slots and its own B in the fused launch, the persistent
    kernel starts at the exchange, B is all-reduced one alpha late.  Same passes, same scores as one GPU
    without the fusion; and the same through cge_b200_score_multi (threads of one process, B summed on the
    host)."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp

    from cge_jl_b200 import divergence as dv
    from util import RTOL, empty_landmark_args

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, False, True, 1, q, True)) for r in range(world)]
    for p in procs:
        p.start()
    out2, iters2, div2, auc2, n_ranks, driver, reg, fused = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert n_ranks == world and driver == 2 and reg == 1 and fused >= 5
    n, edges, ew, vw, comm, emb = _problem(False)
    samples = dv.draw_samples(edges, ew, n, 2000, 42, False, True)
    out1, st1 = dv.wGCL(edges, ew, comm, emb, np.zeros(n), vw, *empty_landmark_args(), False, 42, 2000,
                        False, samples=samples, return_stats=True)
    assert st1.b_fused == 0 and list(st1.iters) == iters2
    assert out1[0] == out2[0] and out1[4] == out2[4]
    np.testing.assert_allclose(out2, out1, rtol=RTOL, atol=1e-15)
    np.testing.assert_allclose(np.array(div2), np.array(list(st1.div)), rtol=RTOL, equal_nan=True)
    if world == 2:
        monkeypatch.setenv("CGE_B200_RT_EXPONENT", "0")
        monkeypatch.setenv("CGE_B200_FUSE_B", "1")
        p, keep = dv.make_problem(edges, ew, comm, emb, np.zeros(n), vw, None, None, None, False, False,
                                  samples, 0, 0, 1)
        out3, st3 = dv.score_multi(p, 2)
        assert st3.b_fused >= 5 and list(st3.iters) == iters2
        np.testing.assert_allclose(out3, out1, rtol=RTOL, atol=1e-15)
        del keep
