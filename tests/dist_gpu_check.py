"""2+ GPU parity check, launched by torchrun (one rank per GPU, NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29517 tests/dist_gpu_check.py
Every rank holds a block of the sample axis of the golden PCA / GMM problems; posteriors and the
lower-bound trajectory must match the single-process reference goldens."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from bayespy_b200 import parallel                                            # noqa: E402
from bayespy_b200.nodes import (GaussianARD, Gamma, SumMultiply, Gaussian, Wishart, Dirichlet,   # noqa: E402
                                Categorical, Mixture)
from bayespy_b200.inference import VB                                        # noqa: E402


def golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))


def main():
    world, rank = parallel.init_from_env()
    g = golden("pca_64x16")
    M, N, K = 64, 96, 16
    n0, n1 = parallel.shard_bounds(N, world, rank)
    X = GaussianARD(0, 1, plates=(1, n1 - n0), shape=(K,), name="X")
    alpha = Gamma(1e-5, 1e-5, plates=(K,), name="alpha")
    C = GaussianARD(0, alpha, plates=(M, 1), shape=(K,), name="C")
    F = SumMultiply("d,d->", X, C)
    tau = Gamma(1e-5, 1e-5, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    Y.observe(g["y"][:, n0:n1])
    C.initialize_from_value(g["C_init"])
    Q = VB(Y, X, C, alpha, tau)
    iters = len(g["L"])
    Q.update(repeat=iters, verbose=False, tol=0)
    np.testing.assert_allclose(Q.L[:iters], g["L"], rtol=1e-8)
    np.testing.assert_allclose(np.asarray(X.u[0]), g["X_u0"][:, n0:n1], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(np.asarray(C.u[0]), g["C_u0"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(np.asarray(tau.u[0]), g["tau_u0"], rtol=1e-8)
    Ls = parallel.allgather_scalar(Q.L[iters - 1])
    assert np.all(Ls == Ls[0]), Ls

    g = golden("gmm_small")
    N, D, K = 300, 3, 5
    n0, n1 = parallel.shard_bounds(N, world, rank)
    al = Dirichlet(1e-5 * np.ones(K), name="alpha")
    Z = Categorical(al, plates=(n1 - n0,), name="Z")
    mu = Gaussian(np.zeros(D), 1e-5 * np.identity(D), plates=(K,), name="mu")
    Lam = Wishart(D, 1e-5 * np.identity(D), plates=(K,), name="Lambda")
    Ym = Mixture(Z, Gaussian, mu, Lam, name="Y")
    Z.initialize_from_value(np.argmax(g["Z_init"][n0:n1], axis=-1))
    Ym.observe(g["y"][n0:n1])
    Q = VB(Ym, mu, Lam, Z, al)
    iters = len(g["L"])
    Q.update(repeat=iters, verbose=False, tol=0)
    np.testing.assert_allclose(Q.L[:iters], g["L"], rtol=1e-8)
    np.testing.assert_allclose(np.asarray(Z.u[0]), g["Z_u0"][n0:n1], rtol=1e-6, atol=1e-10)
    np.testing.assert_allclose(np.asarray(mu.u[0]), g["mu_u0"], rtol=1e-6, atol=1e-9)
    parallel.barrier()
    if rank == 0:
        print("dist_gpu_check OK: world=%d PCA+GMM sharded parity vs reference goldens" % world)


if __name__ == "__main__":
    main()
