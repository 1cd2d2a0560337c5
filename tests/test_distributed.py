"""Plate sharding: the sample axis N split over 2 ranks must reproduce the
single-process result.  CPU version: world_size=2 over gloo, with the oracle backend's
all-reduce hook standing in for ncclAllReduce (host-side sharding logic under test).
GPU version (2 GPUs, NCCL) lives in tests/test_distributed_gpu.py-style runs of bench.py."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, golden


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, model, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import torch
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    from oracle.bpk_ref import RefBackend
    from bayespy_b200 import _bpk, parallel
    be = RefBackend()

    def hook(v):
        t = torch.from_numpy(np.ascontiguousarray(v))
        dist.all_reduce(t)
        return t.numpy()
    be._allreduce_hook = hook
    _bpk._set_backend_for_testing(be)
    parallel.set_world_for_testing(world, rank)
    try:
        if model == "pca":
            from test_models import build_pca
            g = golden("pca_small")
            from bayespy_b200.nodes import GaussianARD, Gamma, SumMultiply
            from bayespy_b200.inference import VB
            M, N, K = 20, 100, 5
            n0, n1 = parallel.shard_bounds(N, world, rank)
            y = g["y"][:, n0:n1]
            X = GaussianARD(0, 1, plates=(1, n1 - n0), shape=(K,), name="X")
            alpha = Gamma(1e-5, 1e-5, plates=(K,), name="alpha")
            C = GaussianARD(0, alpha, plates=(M, 1), shape=(K,), name="C")
            F = SumMultiply("d,d->", X, C)
            tau = Gamma(1e-5, 1e-5, name="tau")
            Y = GaussianARD(F, tau, name="Y")
            Y.observe(y)
            C.initialize_from_value(g["C_init"])
            Q = VB(Y, X, C, alpha, tau)
            iters = len(g["L"])
            Q.update(repeat=iters, verbose=False, tol=0)
            ok = np.allclose(Q.L[:iters], g["L"], rtol=1e-8) and \
                np.allclose(np.asarray(X.u[0]), g["X_u0"][:, n0:n1], rtol=1e-7, atol=1e-9) and \
                np.allclose(np.asarray(C.u[0]), g["C_u0"], rtol=1e-7, atol=1e-9) and \
                np.allclose(np.asarray(tau.u[0]), g["tau_u0"], rtol=1e-8)
            q.put((rank, bool(ok), float(Q.L[iters - 1])))
        else:
            from bayespy_b200.nodes import Gaussian, Wishart, Dirichlet, Categorical, Mixture
            from bayespy_b200.inference import VB
            g = golden("gmm_small")
            N, D, K = 300, 3, 5
            n0, n1 = parallel.shard_bounds(N, world, rank)
            y = g["y"][n0:n1]
            alpha = Dirichlet(1e-5 * np.ones(K), name="alpha")
            Z = Categorical(alpha, plates=(n1 - n0,), name="Z")
            mu = Gaussian(np.zeros(D), 1e-5 * np.identity(D), plates=(K,), name="mu")
            Lambda = Wishart(D, 1e-5 * np.identity(D), plates=(K,), name="Lambda")
            Y = Mixture(Z, Gaussian, mu, Lambda, name="Y")
            Z.initialize_from_value(np.argmax(g["Z_init"][n0:n1], axis=-1))
            Y.observe(y)
            Q = VB(Y, mu, Lambda, Z, alpha)
            iters = len(g["L"])
            Q.update(repeat=iters, verbose=False, tol=0)
            ok = np.allclose(Q.L[:iters], g["L"], rtol=1e-8) and \
                np.allclose(np.asarray(Z.u[0]), g["Z_u0"][n0:n1], rtol=1e-7, atol=1e-10) and \
                np.allclose(np.asarray(mu.u[0]), g["mu_u0"], rtol=1e-7, atol=1e-9)
            q.put((rank, bool(ok), float(Q.L[iters - 1])))
    except Exception as e:                                    # pragma: no cover
        import traceback
        q.put((rank, False, traceback.format_exc()))
    dist.destroy_process_group()


@pytest.mark.parametrize("model", ["pca", "gmm"])
def test_two_rank_gloo_matches_single_process(model):
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, model, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, info in res:
        assert ok, "rank %d: %s" % (rank, info)
    assert res[0][2] == res[1][2]          # bitwise identical bound on both ranks


def test_shard_bounds():
    from bayespy_b200.parallel import shard_bounds
    for n, w in ((10, 3), (10_000_000, 8), (7, 8), (0, 2)):
        blocks = [shard_bounds(n, w, r) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1


def _rdzv_worker(rank, world, port, gloo_port, q):
    """parallel.init_from_env end to end on CPU: NCCL-id file rendezvous, IPC-handle exchange, cleanup."""
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_PORT=str(port),
                      TORCHELASTIC_RUN_ID="pytest-%d" % port)
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % gloo_port, rank=rank, world_size=world)
    from oracle.bpk_ref import RefBackend
    from bayespy_b200 import _bpk, parallel

    class FakeDevice(RefBackend):
        opened = None

        def comm_unique_id(self):
            return bytes([7]) * 128

        def comm_init(self, uid, nranks, r):
            assert uid == bytes([7]) * 128
            super().comm_init(uid, nranks, r)

        def xchg_create(self):
            return bytes([100 + rank]) * 64

        def xchg_open(self, handles, nranks, r):
            FakeDevice.opened = (list(handles), nranks, r)

        def xchg_close(self):
            FakeDevice.opened = None

    be = FakeDevice()

    def hook(v):
        t = torch.from_numpy(np.ascontiguousarray(v))
        dist.all_reduce(t)
        return t.numpy()
    be._allreduce_hook = hook
    _bpk._set_backend_for_testing(be)
    try:
        w, r = parallel.init_from_env(timeout=60.0)
        ok = (w, r) == (world, rank) and parallel._state.get("p2p") is True
        handles, nr, rr = FakeDevice.opened
        ok = ok and nr == world and rr == rank and handles == [bytes([100 + k]) * 64 for k in range(world)]
        parallel.barrier()
        # aligned barrier: every rank leaves within a millisecond of the agreed instant (shared monotonic clock)
        import time
        # (generous bounds: on a loaded CI container three Python processes are descheduled for tens of milliseconds)
        parallel.barrier_aligned(lead_s=0.3)
        t_leave = time.monotonic()
        spread = parallel.allgather_scalar(t_leave)
        ok = ok and float(np.max(spread) - np.min(spread)) < 0.25
        left = [f for f in os.listdir("/tmp") if f.startswith("bpk_rdzv_%d_pytest-%d" % (port, port))]
        q.put((rank, bool(ok), left if rank == 0 else []))
    except Exception:                                         # pragma: no cover
        import traceback
        q.put((rank, False, traceback.format_exc()))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_init_from_env_rendezvous_and_peer_window_exchange(world):
    import multiprocessing as mp
    import time
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port, gloo_port = _free_port(), _free_port()
    procs = [ctx.Process(target=_rdzv_worker, args=(r, world, port, gloo_port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, info in res:
        assert ok, "rank %d: %s" % (rank, info)
    time.sleep(0.2)
    assert not [f for f in os.listdir("/tmp") if f.startswith("bpk_rdzv_%d_pytest-%d" % (port, port))]
