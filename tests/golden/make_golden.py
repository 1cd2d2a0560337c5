"""Generate the golden vectors in tests/golden/*.npz by running the UNMODIFIED
reference (read-only at /root/reference) in the build container.

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box, so its outputs are committed here
as small fixtures; tests compare the oracle (CPU) and the CUDA path (GPU)
against them.  ``h5py`` / ``truncnorm`` are stubbed (tests/golden/_stubs): both
are imported by the reference at module import but never used on this path.
"""
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_stubs"))
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

import numpy as np                                            # noqa: E402
import bayespy                                                # noqa: E402
from bayespy.nodes import (GaussianARD, Gamma, SumMultiply, Gaussian, Wishart, Dirichlet,     # noqa: E402
                           Categorical, Mixture)
from bayespy.inference import VB                              # noqa: E402
from bayespy.utils import misc, linalg, random                # noqa: E402

assert bayespy.__file__.startswith("/root/reference"), bayespy.__file__


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, os.path.getsize(path), "bytes")


def node_state(prefix, node, out):
    # copies: the reference updates node.u in place, a snapshot taken mid-script must not alias it
    for i, u in enumerate(node.u):
        out["%s_u%d" % (prefix, i)] = np.array(u, copy=True)
    for i, p in enumerate(node.phi):
        out["%s_phi%d" % (prefix, i)] = np.array(p, copy=True)
    out["%s_g" % prefix] = np.array(node.g, copy=True)


def quickstart():
    """doc/source/user_guide/quickstart.rst:8-118."""
    np.random.seed(1)
    data = np.random.normal(5, 10, size=(10,))
    mu = GaussianARD(0, 1e-6)
    tau = Gamma(1e-6, 1e-6)
    y = GaussianARD(mu, tau, plates=(10,))
    y.observe(data)
    Q = VB(mu, tau, y)
    Q.update(repeat=20, verbose=False)
    out = dict(data=data, L=Q.L[:Q.iter], iters=Q.iter)
    node_state("mu", mu, out)
    node_state("tau", tau, out)
    save("quickstart", **out)


def pca(name, M, N, K, mask_p=None, iters=8):
    """doc/source/examples/pca.rst:26-66 (no rotation), demos/pca.py:21-71."""
    np.random.seed(1)
    w = np.random.randn(M, 4)
    x = np.random.randn(N, 4)
    y = w @ x.T + 0.1 * np.random.randn(M, N)
    X = GaussianARD(0, 1, plates=(1, N), shape=(K,))
    alpha = Gamma(1e-5, 1e-5, plates=(K,))
    C = GaussianARD(0, alpha, plates=(M, 1), shape=(K,))
    F = SumMultiply('d,d->', X, C)
    tau = Gamma(1e-5, 1e-5)
    Y = GaussianARD(F, tau)
    out = dict(y=y)
    if mask_p is not None:
        mask = random.mask(M, N, p=mask_p)
        Y.observe(y, mask=mask)
        out["mask"] = mask
    else:
        Y.observe(y)
    C.initialize_from_random()
    out["C_init"] = np.array(C.u[0], copy=True)
    Q = VB(Y, X, C, alpha, tau)
    Q.update(repeat=iters, verbose=False, tol=0)
    out["L"] = Q.L[:Q.iter]
    out["l_Y"], out["l_X"], out["l_C"] = Q.l[Y][:Q.iter], Q.l[X][:Q.iter], Q.l[C][:Q.iter]
    out["l_alpha"], out["l_tau"] = Q.l[alpha][:Q.iter], Q.l[tau][:Q.iter]
    for nm, node in (("X", X), ("C", C), ("alpha", alpha), ("tau", tau)):
        node_state(nm, node, out)
    # messages arriving at the stochastic parents (reference-format composite messages)
    mC = C._message_from_children()
    out["msgC0"], out["msgC1"] = mC[0], mC[1]
    mt = tau._message_from_children()
    out["msgtau0"], out["msgtau1"] = np.asarray(mt[0]), np.asarray(mt[1])
    save(name, **out)


def linalg_vectors():
    """bayespy/utils/linalg.py:31-223 on random SPD stacks."""
    rng = np.random.RandomState(7)
    out = {}
    for D in (1, 3, 8, 16, 33):
        A = rng.randn(5, 2, D, D)
        A = A @ np.swapaxes(A, -1, -2) + D * np.identity(D)
        b = rng.randn(5, 2, D)
        B = rng.randn(1, 2, D, 4)
        U = linalg.chol(A)
        out["A%d" % D], out["b%d" % D], out["B%d" % D] = A, b, B
        out["U%d" % D] = np.triu(U)
        out["solve%d" % D] = linalg.chol_solve(U, b)
        out["solveM%d" % D] = linalg.chol_solve(U, B * np.ones((5, 1, 1, 1)), matrix=True)
        out["inv%d" % D] = linalg.chol_inv(U)
        out["logdet%d" % D] = linalg.chol_logdet(U)
    save("linalg", **out)


def summul_vectors():
    """bayespy/utils/misc.py:805-945."""
    rng = np.random.RandomState(11)
    a = rng.randn(4, 1, 5)
    b = rng.randn(3, 5)
    c = rng.randn(5)
    out = dict(a=a, b=b, c=c)
    out["r0"] = misc.sum_multiply(a, b, c)
    out["r1"] = misc.sum_multiply(a, b, c, axis=-1)
    out["r2"] = misc.sum_multiply(a, b, c, axis=(0, 2), keepdims=True)
    out["r3"] = misc.sum_multiply(a, b, axis=[1], sumaxis=False)
    out["r4"] = misc.sum_product(a, b, c, axes_to_keep=[0, 2])
    out["r5"] = misc.sum_product(a, b, axes_to_sum=[-2], keepdims=True)
    out["p0"] = misc.sum_multiply_to_plates(a, b, to_plates=(3, 1), from_plates=(4, 3, 5))
    out["p1"] = misc.sum_multiply_to_plates(a, b, to_plates=(1,), from_plates=(4, 3, 5))
    out["p2"] = misc.sum_multiply_to_plates(c, to_plates=(), from_plates=(6, 5))
    A = rng.randn(4, 3, 2, 2)
    out["A"] = A
    out["p3"] = misc.sum_multiply_to_plates(A, to_plates=(3,), from_plates=(4, 3), ndim=2)
    save("summul", **out)


def distribution_vectors():
    """compute_moments_and_cgf of the scalar/simplex families on random phi."""
    from bayespy.inference.vmp.nodes.gamma import GammaDistribution
    from bayespy.inference.vmp.nodes.wishart import WishartDistribution
    from bayespy.inference.vmp.nodes.dirichlet import DirichletDistribution
    from bayespy.inference.vmp.nodes.categorical import CategoricalDistribution
    from bayespy.inference.vmp.nodes.gaussian import GaussianARDDistribution
    rng = np.random.RandomState(3)
    out = {}
    phi = [-rng.gamma(2.0, 1.0, size=(7, 3)), rng.gamma(3.0, 1.0, size=(7, 3)) + 0.01]
    (u, g) = GammaDistribution().compute_moments_and_cgf(phi)
    out.update(gam_phi0=phi[0], gam_phi1=phi[1], gam_u0=u[0], gam_u1=u[1], gam_g=g)
    D = 4
    V = rng.randn(6, D, D)
    V = V @ np.swapaxes(V, -1, -2) + np.identity(D)
    phi = [-0.5 * V, 0.5 * (D + rng.gamma(2.0, 2.0, size=(6,)))]
    (u, g) = WishartDistribution().compute_moments_and_cgf(phi)
    out.update(wis_phi0=phi[0], wis_phi1=phi[1], wis_u0=u[0], wis_u1=u[1], wis_g=g)
    phi = [rng.gamma(1.0, 2.0, size=(5, 6)) + 1e-3]
    (u, g) = DirichletDistribution().compute_moments_and_cgf(phi)
    out.update(dir_phi0=phi[0], dir_u0=u[0], dir_g=g)
    phi = [5 * rng.randn(9, 6)]
    (u, g) = CategoricalDistribution(6).compute_moments_and_cgf(phi)
    out.update(cat_phi0=phi[0], cat_u0=u[0], cat_g=g)
    K = 5
    L = rng.randn(8, K, K)
    L = L @ np.swapaxes(L, -1, -2) + np.identity(K)
    phi = [rng.randn(8, K), -0.5 * L]
    (u, g) = GaussianARDDistribution((K,)).compute_moments_and_cgf(phi)
    out.update(gau_phi0=phi[0], gau_phi1=phi[1], gau_u0=u[0], gau_u1=u[1], gau_g=g)
    phi = [rng.randn(8, K), -0.5 * L[:1]]
    (u, g) = GaussianARDDistribution((K,)).compute_moments_and_cgf(phi)
    out.update(gaus_phi0=phi[0], gaus_phi1=phi[1], gaus_u0=u[0], gaus_u1=u[1], gaus_g=g)
    save("distributions", **out)


def gmm(name, N, D, K, iters=10):
    """doc/source/examples/gmm.rst:71-98 on blobs (demos/stochastic_inference.py:49-54 pattern)."""
    np.random.seed(1)
    means = 5 * np.random.randn(K, D)
    z = np.random.randint(0, K, size=N)
    y = means[z] + np.random.randn(N, D)
    alpha = Dirichlet(1e-5 * np.ones(K))
    Z = Categorical(alpha, plates=(N,))
    mu = Gaussian(np.zeros(D), 1e-5 * np.identity(D), plates=(K,))
    Lambda = Wishart(D, 1e-5 * np.identity(D), plates=(K,))
    Y = Mixture(Z, Gaussian, mu, Lambda)
    Z.initialize_from_random()
    out = dict(y=y, Z_init=np.array(Z.u[0], copy=True))
    Y.observe(y)
    Q = VB(Y, mu, Lambda, Z, alpha)
    Q.update(repeat=iters, verbose=False, tol=0)
    out["L"] = Q.L[:Q.iter]
    for nm, node in (("Z", Z), ("mu", mu), ("Lambda", Lambda), ("alpha", alpha)):
        node_state(nm, node, out)
        out["l_" + nm] = Q.l[node][:Q.iter]
    out["l_Y"] = Q.l[Y][:Q.iter]
    save(name, **out)


def gmm_doc():
    """doc/source/examples/gmm.rst:8-118: literal doctest trajectory (-1.402345e+03 ... -8.888464e+02)."""
    np.random.seed(1)
    y0 = np.random.multivariate_normal([0, 0], [[2, 0], [0, 0.1]], size=50)
    y1 = np.random.multivariate_normal([0, 0], [[0.1, 0], [0, 2]], size=50)
    y2 = np.random.multivariate_normal([2, 2], [[2, -1.5], [-1.5, 2]], size=50)
    y3 = np.random.multivariate_normal([-2, -2], [[0.5, 0], [0, 0.5]], size=50)
    y = np.vstack([y0, y1, y2, y3])
    N, D, K = 200, 2, 10
    alpha = Dirichlet(1e-5 * np.ones(K), name='alpha')
    Z = Categorical(alpha, plates=(N,), name='z')
    mu = Gaussian(np.zeros(D), 1e-5 * np.identity(D), plates=(K,), name='mu')
    Lambda = Wishart(D, 1e-5 * np.identity(D), plates=(K,), name='Lambda')
    Y = Mixture(Z, Gaussian, mu, Lambda, name='Y')
    Z.initialize_from_random()
    Zi = np.array(Z.u[0], copy=True)
    Q = VB(Y, mu, Lambda, Z, alpha)
    Y.observe(y)
    Q.update(repeat=1000, verbose=False)
    save("gmm_doc", y=y, Z_init=Zi, L=Q.L[:Q.iter], iters=Q.iter)


def block_banded_vectors():
    """utils/linalg.py:468-575 block_banded_solve on random SPD block-tridiagonal systems
    (same construction as utils/tests/test_linalg.py:112-182)."""
    out = {}
    rs = np.random.RandomState(11)
    for tag, (N, D) in dict(a=(7, 3), b=(40, 5), c=(3, 1), d=(12, 16)).items():
        # SPD by construction: diagonally dominant blocks
        R = rs.randn(N * D, N * D)
        full = R @ R.T + N * D * np.identity(N * D)
        A = np.stack([full[n * D:(n + 1) * D, n * D:(n + 1) * D] for n in range(N)])
        B = np.stack([full[n * D:(n + 1) * D, (n + 1) * D:(n + 2) * D] for n in range(N - 1)]) \
            if N > 1 else np.zeros((0, D, D))
        y = rs.randn(N, D)
        V, C, x, ldet = linalg.block_banded_solve(A, B, y)
        out.update({"A_" + tag: A, "B_" + tag: B, "y_" + tag: y, "V_" + tag: V, "C_" + tag: C, "x_" + tag: x,
                    "ldet_" + tag: np.asarray(ldet)})
    save("block_banded", **out)


def lssm(name, M, N, D, mask_p=None, iters=6):
    """doc/source/examples/lssm.rst:45-181 scaled down: X = GaussianMarkovChain(0, 1e-3 I, A, 1), F = Dot(C, X)."""
    from bayespy.nodes import GaussianMarkovChain, Dot
    np.random.seed(5)
    # two noisy oscillators + random walk (lssm.rst:153-175 pattern)
    w = 0.3
    a = np.array([[np.cos(w), -np.sin(w), 0], [np.sin(w), np.cos(w), 0], [0, 0, 1.0]])
    x = np.empty((N, 3))
    x[0] = 10 * np.random.randn(3)
    for n in range(N - 1):
        x[n + 1] = a @ x[n] + [1, 1, 10] * np.random.randn(3) * 0.1
    c = np.random.randn(M, 3)
    y = x @ c.T + 0.5 * np.random.randn(N, M)
    y = y.T                                              # (M, N)
    alpha = Gamma(1e-5, 1e-5, plates=(D,), name="alpha")
    A = GaussianARD(0, alpha, shape=(D,), plates=(D,), name="A")
    X = GaussianMarkovChain(np.zeros(D), 1e-3 * np.identity(D), A, np.ones(D), n=N, name="X")
    gamma = Gamma(1e-5, 1e-5, plates=(D,), name="gamma")
    C = GaussianARD(0, gamma, shape=(D,), plates=(M, 1), name="C")
    F = Dot(C, X, name="F")
    C_init = np.random.RandomState(3).randn(M, 1, D)
    C.initialize_from_value(C_init)
    tau = Gamma(1e-5, 1e-5, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    mask = True
    if mask_p is not None:
        mask = random.mask(M, N, p=mask_p)
        Y.observe(y, mask=mask)
    else:
        Y.observe(y)
    Q = VB(X, C, gamma, A, alpha, tau, Y)
    Q.update(repeat=iters, verbose=False, tol=0)
    out = dict(y=y, C_init=C_init, L=Q.L[:iters], mask=np.asarray(mask))
    for nm, node in (("X", X), ("C", C), ("gamma", gamma), ("A", A), ("alpha", alpha), ("tau", tau)):
        node_state(nm, node, out)
    for node in Q.model:
        out["l_" + node.name] = Q.l[node][:iters]
    save(name, **out)


def pca_rotated(name="pca_rotated", M=10, N=60, D=4, iters=6):
    """doc/source/examples/pca.rst:26-112: Bayesian PCA with the rotation parameter expansion as VB callback
    (inference/vmp/transformations.py:23-222 RotationOptimizer, :376-1110 RotateGaussianARD)."""
    from bayespy.inference.vmp.transformations import RotateGaussianARD, RotationOptimizer
    np.random.seed(1)
    c = np.random.randn(M, 2)
    x = np.random.randn(2, N)
    y = np.dot(c, x) + 0.1 * np.random.randn(M, N)
    X = GaussianARD(0, 1, shape=(D,), plates=(1, N), name="X")
    alpha = Gamma(1e-5, 1e-5, plates=(D,), name="alpha")
    C = GaussianARD(0, alpha, shape=(D,), plates=(M, 1), name="C")
    F = SumMultiply("d,d->", X, C)
    tau = Gamma(1e-5, 1e-5, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    Y.observe(y)
    C_init = np.random.RandomState(7).randn(M, 1, D)
    C.initialize_from_value(C_init)
    Q = VB(Y, X, C, alpha, tau)
    rot_X = RotateGaussianARD(X)
    rot_C = RotateGaussianARD(C, alpha)
    R = RotationOptimizer(rot_X, rot_C, D)
    # one explicit rotation from a known state, to pin the cost function and its gradient
    Q.update(repeat=2, verbose=False, tol=0)
    rot_X.setup(); rot_C.setup()
    Rtest = np.identity(D) + 0.1 * np.random.RandomState(3).randn(D, D)
    bX, dbX = rot_X.bound(Rtest)
    bC, dbC = rot_C.bound(np.linalg.inv(Rtest).T)
    Q.set_callback(R.rotate)
    Q.update(repeat=iters, verbose=False, tol=0)
    out = dict(y=y, C_init=C_init, L=Q.L[:iters + 2], Rtest=Rtest, bX=np.asarray(bX), dbX=dbX, bC=np.asarray(bC), dbC=dbC)
    for nm, node in (("X", X), ("C", C), ("alpha", alpha), ("tau", tau)):
        node_state(nm, node, out)
    save(name, **out)


def summultiply_nodes():
    """nodes/dot.py:19-633 SumMultiply: moments and messages for several key patterns and plate layouts
    (the cases of nodes/tests/test_dot.py), as computed by the reference."""
    out = {}
    rs = np.random.RandomState(21)

    def gauss(shape, plates, nm):
        n = len(shape)
        x = GaussianARD(rs.randn(*(plates + shape)), 1 + rs.rand(*(plates + shape)), shape=shape, plates=plates, name=nm)
        return x
    cases = {
        "dot":      ("i,i", [((4,), (3, 1)), ((4,), (1, 5))]),
        "matvec":   ("ij,j->i", [((3, 4), (2,)), ((4,), (2,))]),
        "outer":    ("i,j->ij", [((3,), (5,)), ((2,), (1,))]),
        "triple":   ("i,i,i->", [((3,), (2, 1)), ((3,), (1, 4)), ((3,), ())]),
        "keepdim":  ("ij,ik->jk", [((2, 3), ()), ((2, 4), (3,))]),
    }
    for tag, (spec, ins) in cases.items():
        nodes = [gauss(sh, pl, "%s_x%d" % (tag, i)) for i, (sh, pl) in enumerate(ins)]
        F = SumMultiply(spec, *nodes)
        u = F.get_moments()
        out[tag + "_u0"] = np.asarray(u[0])
        out[tag + "_u1"] = np.asarray(u[1])
        out[tag + "_plates"] = np.asarray(F.plates, dtype=np.int64)
        # a child so that messages flow: Y ~ GaussianARD(F, tau) observed
        tau = 0.7 + rs.rand()
        ndim = len(F.dims[0])
        Y = GaussianARD(F, tau, ndim=ndim)
        yv = rs.randn(*(tuple(F.plates) + tuple(F.dims[0])))
        Y.observe(yv)
        out[tag + "_y"] = yv
        out[tag + "_tau"] = np.asarray(tau)
        for i, nd in enumerate(nodes):
            out["%s_in%d_mu" % (tag, i)] = np.asarray(nd.parents[0].get_moments()[0]) / np.asarray(nd.parents[0].get_moments()[2]) \
                if False else np.asarray(nd.u[0])
            out["%s_in%d_u1" % (tag, i)] = np.asarray(nd.u[1])
            m = F._message_to_parent(i)
            out["%s_m%d_0" % (tag, i)] = np.asarray(m[0])
            out["%s_m%d_1" % (tag, i)] = np.asarray(m[1])
    save("summultiply_nodes", **out)


def mixture_ard(name="mixture_ard", N=120, K=3, iters=6):
    """Mixture of scalar GaussianARD components (mixture.py:26-488 with a mixed class other than Gaussian)."""
    np.random.seed(8)
    y = np.concatenate([np.random.randn(N // 3) - 4, 0.5 * np.random.randn(N // 3), 2 * np.random.randn(N - 2 * (N // 3)) + 5])
    alpha = Dirichlet(1e-3 * np.ones(K), name="alpha")
    Z = Categorical(alpha, plates=(N,), name="Z")
    mu = GaussianARD(0, 1e-3, plates=(K,), name="mu")
    tau = Gamma(1e-3, 1e-3, plates=(K,), name="tau")
    Y = Mixture(Z, GaussianARD, mu, tau, name="Y")
    Z_init = np.random.RandomState(2).dirichlet(np.ones(K), size=N)
    Z.initialize_from_value(np.argmax(Z_init, axis=-1))
    Y.observe(y)
    Q = VB(Y, mu, tau, Z, alpha)
    Q.update(repeat=iters, verbose=False, tol=0)
    out = dict(y=y, Z_init=Z_init, L=Q.L[:iters])
    for nm, node in (("Z", Z), ("mu", mu), ("tau", tau), ("alpha", alpha)):
        node_state(nm, node, out)
    save(name, **out)


def lssm_plated(name="lssm_plated", M=5, N=30, D=3, P=2, iters=5):
    """Two independent chains (plates (P,)) sharing the dynamics A and the loadings C
    (gaussian_markov_chain.py:660-927 with plates)."""
    from bayespy.nodes import GaussianMarkovChain, Dot
    rs = np.random.RandomState(12)
    y = rs.randn(M, P, N).cumsum(axis=-1) * 0.3 + rs.randn(M, P, N)
    alpha = Gamma(1e-5, 1e-5, plates=(D,), name="alpha")
    A = GaussianARD(0, alpha, shape=(D,), plates=(D,), name="A")
    mu0 = rs.randn(P, D)
    X = GaussianMarkovChain(mu0, 1e-3 * np.identity(D), A, np.ones(D), n=N, name="X")
    assert X.plates == (P,)
    gamma = Gamma(1e-5, 1e-5, plates=(D,), name="gamma")
    C = GaussianARD(0, gamma, shape=(D,), plates=(M, 1, 1), name="C")
    F = Dot(C, X, name="F")
    assert F.plates == (M, P, N)
    C_init = rs.randn(M, 1, 1, D)
    C.initialize_from_value(C_init)
    tau = Gamma(1e-5, 1e-5, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    Y.observe(y)
    Q = VB(X, C, gamma, A, alpha, tau, Y)
    Q.update(repeat=iters, verbose=False, tol=0)
    out = dict(y=y, C_init=C_init, mu0=mu0, L=Q.L[:iters])
    for nm, node in (("X", X), ("C", C), ("A", A), ("alpha", alpha), ("tau", tau)):
        node_state(nm, node, out)
    save(name, **out)


def lssm_varying(name="lssm_varying", M=5, N=12, D=2, iters=4):
    """One chain with a transition matrix and an innovation precision PER STEP: A and nu with plates (N-1, D)
    (gaussian_markov_chain.py:660-706, 840-880; the chain length is inferred from the parents)."""
    from bayespy.nodes import GaussianMarkovChain, Dot
    rs = np.random.RandomState(21)
    y = rs.randn(M, N).cumsum(axis=-1) * 0.3 + rs.randn(M, N)
    alpha = Gamma(1e-5, 1e-5, plates=(D,), name="alpha")
    A = GaussianARD(0, alpha, shape=(D,), plates=(N - 1, D), name="A")
    A_init = 0.5 * rs.randn(N - 1, D, D)
    A.initialize_from_value(A_init)
    nu = Gamma(1e-3, 1e-3, plates=(N - 1, D), name="nu")
    X = GaussianMarkovChain(np.zeros(D), 1e-3 * np.identity(D), A, nu, name="X")
    assert X.plates == () and X.dims[0] == (N, D)
    gamma = Gamma(1e-5, 1e-5, plates=(D,), name="gamma")
    C = GaussianARD(0, gamma, shape=(D,), plates=(M, 1), name="C")
    F = Dot(C, X, name="F")
    C_init = rs.randn(M, 1, D)
    C.initialize_from_value(C_init)
    tau = Gamma(1e-5, 1e-5, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    Y.observe(y)
    Q = VB(X, C, gamma, A, alpha, nu, tau, Y)
    Q.update(repeat=iters, verbose=False, tol=0)
    out = dict(y=y, C_init=C_init, A_init=A_init, L=Q.L[:iters])
    for nm, node in (("X", X), ("C", C), ("A", A), ("alpha", alpha), ("nu", nu), ("tau", tau)):
        node_state(nm, node, out)
    save(name, **out)


def lssm_mixing(name="lssm_mixing", M=5, N=14, D=2, K=3, iters=5):
    """VaryingGaussianMarkovChain (gaussian_markov_chain.py:930-1452): A_n = sum_k s_nk B_k with B a GaussianARD of
    shape (D, K), plates (D,) and the weights S plated over time; observed through Dot."""
    from bayespy.nodes import VaryingGaussianMarkovChain, Dot
    rs = np.random.RandomState(33)
    y = rs.randn(M, N).cumsum(axis=-1) * 0.3 + rs.randn(M, N)
    beta = Gamma(1e-3, 1e-3, plates=(K,), name="beta")
    B = GaussianARD(0, beta, shape=(D, K), plates=(D,), name="B")
    B_init = 0.5 * rs.randn(D, D, K)
    B.initialize_from_value(B_init)
    S = GaussianARD(0, 1, shape=(K,), plates=(N - 1,), name="S")
    S_init = rs.randn(N - 1, K)
    S.initialize_from_value(S_init)
    nu = np.array([1.0, 2.5])[:D]            # the reference has no message to the innovation precision of this node
    X = VaryingGaussianMarkovChain(np.zeros(D), 1e-3 * np.identity(D), B, S, nu, name="X")
    assert X.plates == () and X.dims[0] == (N, D)
    gamma = Gamma(1e-5, 1e-5, plates=(D,), name="gamma")
    C = GaussianARD(0, gamma, shape=(D,), plates=(M, 1), name="C")
    F = Dot(C, X, name="F")
    C_init = rs.randn(M, 1, D)
    C.initialize_from_value(C_init)
    tau = Gamma(1e-5, 1e-5, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    Y.observe(y)
    Q = VB(X, C, gamma, B, beta, S, tau, Y)
    Q.update(repeat=iters, verbose=False, tol=0)
    out = dict(y=y, C_init=C_init, B_init=B_init, S_init=S_init, nu=nu, L=Q.L[:iters])
    for nm, node in (("X", X), ("C", C), ("B", B), ("beta", beta), ("S", S), ("tau", tau)):
        node_state(nm, node, out)
    for node in Q.model:
        out["l_" + node.name] = Q.l[node][:iters]
    save(name, **out)


def lssm_switching(name="lssm_switching", M=5, N=14, D=2, K=3, iters=5):
    """SwitchingGaussianMarkovChain (gaussian_markov_chain.py:1454-1985): A_n = B_{z_n} with z_n categorical over K
    transition matrices; B a GaussianARD of shape (D,), plates (K, D)."""
    from bayespy.nodes import SwitchingGaussianMarkovChain, Dot
    rs = np.random.RandomState(34)
    y = rs.randn(M, N).cumsum(axis=-1) * 0.3 + rs.randn(M, N)
    beta = Gamma(1e-3, 1e-3, plates=(K, 1, 1), name="beta")
    B = GaussianARD(0, beta, shape=(D,), plates=(K, D), name="B")     # (..., K, D) plates, D-dimensional rows (:1878-1880)
    B_init = 0.5 * rs.randn(K, D, D)
    B.initialize_from_value(B_init)
    pi = Dirichlet(np.ones(K), name="pi")
    Z = Categorical(pi, plates=(N - 1,), name="Z")
    Z_init = rs.randint(0, K, size=N - 1)
    Z.initialize_from_value(Z_init)
    nu = np.array([1.0, 2.5])[:D]
    X = SwitchingGaussianMarkovChain(np.zeros(D), 1e-3 * np.identity(D), B, Z, nu, name="X")
    assert X.plates == () and X.dims[0] == (N, D)
    gamma = Gamma(1e-5, 1e-5, plates=(D,), name="gamma")
    C = GaussianARD(0, gamma, shape=(D,), plates=(M, 1), name="C")
    F = Dot(C, X, name="F")
    C_init = rs.randn(M, 1, D)
    C.initialize_from_value(C_init)
    tau = Gamma(1e-5, 1e-5, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    Y.observe(y)
    Q = VB(X, C, gamma, B, beta, Z, pi, tau, Y)
    Q.update(repeat=iters, verbose=False, tol=0)
    out = dict(y=y, C_init=C_init, B_init=B_init, Z_init=Z_init, nu=nu, L=Q.L[:iters])
    for nm, node in (("X", X), ("C", C), ("B", B), ("beta", beta), ("Z", Z), ("pi", pi), ("tau", tau)):
        node_state(nm, node, out)
    for node in Q.model:
        out["l_" + node.name] = Q.l[node][:iters]
    save(name, **out)


def pca_gradients(name="pca_gradients", M=8, N=30, D=3):
    """Gradient-based learning on the PCA model (vmp.py:402-662, expfamily.py:260-340, gaussian.py:824-890,
    gamma.py:183-211): Riemannian and Euclidean gradients, a natural-gradient step, Riemannian conjugate gradient with
    collapsed nodes, plain conjugate gradient, pattern search."""
    rs = np.random.RandomState(11)
    y = rs.randn(M, 2) @ rs.randn(2, N) + 0.1 * rs.randn(M, N)
    X = GaussianARD(0, 1, shape=(D,), plates=(1, N), name="X")
    alpha = Gamma(1e-3, 1e-3, plates=(D,), name="alpha")
    C = GaussianARD(0, alpha, shape=(D,), plates=(M, 1), name="C")
    F = SumMultiply("d,d->", X, C)
    tau = Gamma(1e-3, 1e-3, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    Y.observe(y)
    C_init = rs.randn(M, 1, D)
    C.initialize_from_value(C_init)
    Q = VB(Y, X, C, alpha, tau)
    Q.update(repeat=2, verbose=False, tol=0)
    out = dict(y=y, C_init=C_init)
    rg, g = Q.get_gradients(C, tau, X, euclidian=True)
    for nm, r_, g_ in zip(("C", "tau", "X"), rg, g):
        for i in range(2):
            out["rg_%s_%d" % (nm, i)] = np.array(r_[i])
            out["g_%s_%d" % (nm, i)] = np.array(g_[i])
    out["dot"] = Q.dot(rg, g)
    Q.gradient_step(C, tau, scale=0.4)
    out["L_after_step"] = Q.compute_lowerbound()
    node_state("step_C", C, out)
    node_state("step_tau", tau, out)
    Q.optimize(C, tau, maxiter=6, collapsed=[X, alpha], verbose=False, tol=0)
    out["L_opt1"] = Q.L[:Q.iter].copy()
    Q.optimize(C, X, maxiter=4, riemannian=False, verbose=False, tol=0)
    out["L_opt2"] = Q.L[:Q.iter].copy()
    Q.optimize(C, tau, maxiter=3, method="gradient", verbose=False, tol=0)
    out["L_opt3"] = Q.L[:Q.iter].copy()
    Q.pattern_search(C, tau, collapsed=[X, alpha])
    Q.pattern_search(C, X)
    out["L"] = Q.L[:Q.iter].copy()
    for nm, node in (("X", X), ("C", C), ("alpha", alpha), ("tau", tau)):
        node_state(nm, node, out)
    save(name, **out)


def svi_mixture(name="svi_mixture", N=400, N_batch=40, K=4, D=2, steps=12):
    """Stochastic variational inference (demos/stochastic_inference.py:93-141): a mini-batch Mixture whose Categorical
    carries ``plates_multiplier=(N / N_batch,)``, local update of Z, then a natural-gradient step of the global nodes
    with a decaying step length."""
    rs = np.random.RandomState(8)
    means = 4 * rs.randn(3, D)
    data = means[rs.randint(0, 3, size=N)] + rs.randn(N, D)
    mu = Gaussian(np.zeros(D), np.identity(D), plates=(K,), name="means")
    alpha = Dirichlet(np.ones(K), name="class probabilities")
    Z = Categorical(alpha, plates=(N_batch,), plates_multiplier=(N / N_batch,), name="classes")
    Y = Mixture(Z, Gaussian, mu, np.identity(D), name="observations")
    mu_init = rs.randn(K, D)
    mu.initialize_from_value(mu_init)
    Q = VB(Y, Z, mu, alpha)
    Q.ignore_bound_checks = True
    subsets = np.array([rs.choice(N, N_batch) for _ in range(steps)])
    mus, alphas, Ls = [], [], []
    for n in range(steps):
        Y.observe(data[subsets[n], :])
        Q.update(Z, verbose=False)
        step = (n + 1) ** (-0.7)
        Q.gradient_step(mu, alpha, scale=step)
        mus.append(np.array(mu.u[0], copy=True))
        alphas.append(np.array(alpha.u[0], copy=True))
        Ls.append(Q.compute_lowerbound())
    save(name, data=data, mu_init=mu_init, subsets=subsets, mus=np.array(mus), alphas=np.array(alphas), Ls=np.array(Ls),
         L=Q.L[:Q.iter].copy())


def lda(name="lda_small", n_documents=5, n_words=600, n_vocabulary=20, n_topics=3, iters=6, svi_steps=6, subset_size=100):
    """doc/source/examples/lda.rst:84-135 (batch VB) and :180-262 (stochastic VI with ``plates_multiplier``) scaled down:
    Dirichlet / Categorical / Gate and a constant categorical index node."""
    from bayespy import nodes
    from bayespy.inference.vmp.nodes.categorical import CategoricalMoments
    rs = np.random.RandomState(12)
    word_documents = rs.randint(0, n_documents, size=n_words)
    true_topic = rs.dirichlet(0.3 * np.ones(n_topics), size=n_documents)
    true_word = rs.dirichlet(0.2 * np.ones(n_vocabulary), size=n_topics)
    corpus = np.array([rs.choice(n_vocabulary, p=true_word[rs.choice(n_topics, p=true_topic[d])]) for d in word_documents])
    out = dict(word_documents=word_documents, corpus=corpus)
    p_topic = nodes.Dirichlet(np.ones(n_topics), plates=(n_documents,), name="p_topic")
    p_word = nodes.Dirichlet(np.ones(n_vocabulary), plates=(n_topics,), name="p_word")
    document_indices = nodes.Constant(CategoricalMoments(n_documents), word_documents, name="document_indices")
    topics = nodes.Categorical(nodes.Gate(document_indices, p_topic), plates=(len(corpus),), name="topics")
    words = nodes.Categorical(nodes.Gate(topics, p_word), name="words")
    words.observe(corpus)
    np.random.seed(5)
    p_topic.initialize_from_random()
    p_word.initialize_from_random()
    out["p_topic_init"] = np.exp(np.array(p_topic.u[0], copy=True))
    out["p_word_init"] = np.exp(np.array(p_word.u[0], copy=True))
    Q = VB(words, topics, p_word, p_topic, document_indices)
    Q.update(repeat=iters, verbose=False, tol=0)
    out["L"] = Q.L[:iters].copy()
    for nm, node in (("p_topic", p_topic), ("p_word", p_word), ("topics", topics)):
        node_state(nm, node, out)
    # stochastic variational inference
    mult = n_words / subset_size
    p_topic = nodes.Dirichlet(np.ones(n_topics), plates=(n_documents,), name="p_topic")
    p_word = nodes.Dirichlet(np.ones(n_vocabulary), plates=(n_topics,), name="p_word")
    document_indices = nodes.Constant(CategoricalMoments(n_documents), word_documents[:subset_size], name="document_indices")
    topics = nodes.Categorical(nodes.Gate(document_indices, p_topic), plates=(subset_size,), plates_multiplier=(mult,),
                               name="topics")
    words = nodes.Categorical(nodes.Gate(topics, p_word), name="words")
    p_topic.initialize_from_value(out["p_topic_init"])
    p_word.initialize_from_value(out["p_word_init"])
    Q = VB(words, topics, p_word, p_topic, document_indices)
    Q.ignore_bound_checks = True
    subsets = np.array([rs.choice(n_words, subset_size) for _ in range(svi_steps)])
    pt, pw = [], []
    for n in range(svi_steps):
        Q["words"].observe(corpus[subsets[n]])
        Q["document_indices"].set_value(word_documents[subsets[n]])
        Q.update("topics", verbose=False)
        Q.gradient_step("p_topic", "p_word", scale=(n + 1) ** (-0.7))
        pt.append(np.array(p_topic.u[0], copy=True))
        pw.append(np.array(p_word.u[0], copy=True))
    out.update(subsets=subsets, svi_p_topic=np.array(pt), svi_p_word=np.array(pw), svi_L=Q.L[:Q.iter].copy())
    save(name, **out)


def advanced_guide(name="advanced_guide"):
    """doc/source/user_guide/advanced.rst on the user guide's PCA model: deterministic annealing (:219-224) and
    stochastic variational inference with ``plates_multiplier=(1, 20)`` on X (:276-313)."""
    from bayespy.nodes import Dot
    rs = np.random.RandomState(17)
    D = 3
    c = rs.randn(10, 2)
    x = rs.randn(2, 100)
    data = np.dot(c, x) + 0.1 * rs.randn(10, 100)

    def model(nx, mult=None):
        kw = {} if mult is None else dict(plates_multiplier=mult)
        X = GaussianARD(0, 1, shape=(D,), plates=(1, nx), name="X", **kw)
        alpha = Gamma(1e-3, 1e-3, plates=(D,), name="alpha")
        C = GaussianARD(0, alpha, shape=(D,), plates=(10, 1), name="C")
        F = Dot(C, X)
        tau = Gamma(1e-3, 1e-3, name="tau")
        Y = GaussianARD(F, tau, name="Y")
        return X, alpha, C, tau, Y
    out = dict(data=data)
    # deterministic annealing
    X, alpha, C, tau, Y = model(100)
    Y.observe(data)
    Q = VB(Y, C, X, alpha, tau)
    X_init = rs.randn(1, 100, D)
    X.initialize_from_parameters(X_init, 10)
    beta, betas = 0.1, []
    while beta < 1.0:
        beta = min(beta * 1.5, 1.0)
        Q.set_annealing(beta)
        Q.update(repeat=100, tol=1e-4, verbose=False)
        betas.append((beta, Q.iter))
    out.update(X_init=X_init, anneal_L=Q.L[:Q.iter].copy(), anneal_schedule=np.array(betas))
    for nm, node in (("C", C), ("tau", tau), ("alpha", alpha)):
        node_state("anneal_" + nm, node, out)
    # stochastic variational inference
    X, alpha, C, tau, Y = model(5, mult=(1, 20))
    Q = VB(Y, C, X, alpha, tau)
    C_init = rs.randn(10, 1, D)
    C.initialize_from_value(C_init)
    Q.ignore_bound_checks = True
    steps = 15
    subsets = np.array([rs.choice(100, 5) for _ in range(steps)])
    Cs, taus, alphas = [], [], []
    for n in range(steps):
        Y.observe(data[:, subsets[n]])
        Q.update(X, verbose=False)
        Q.gradient_step(C, alpha, tau, scale=(n + 2.0) ** (-0.7))
        Cs.append(np.array(C.u[0], copy=True))
        taus.append(np.array(tau.u[0], copy=True))
        alphas.append(np.array(alpha.u[0], copy=True))
    out.update(C_init=C_init, subsets=subsets, svi_C=np.array(Cs), svi_tau=np.array(taus), svi_alpha=np.array(alphas),
               svi_L=Q.L[:Q.iter].copy())
    save(name, **out)


def sliced_nodes(name="slice"):
    """Basic slicing of plates (node.py:761-763, :868-1160): a plated latent mean observed through several slices —
    ranges with steps, an integer, a new axis, an ellipsis, a negative step."""
    rs = np.random.RandomState(23)
    mu = GaussianARD(0.5, 1e-2, plates=(6, 4), name="mu")
    vec = GaussianARD(0, 1e-1, shape=(2,), plates=(5,), name="vec")
    obs = {}
    # (negative steps are left out: the reference re-normalises its stored slices and ends up with empty plates there)
    children = [("a", mu[1:4, ::2], 2.0), ("b", mu[0], 1.5), ("c", mu[None, 5, 1:3], 3.0), ("d", mu[..., -1], 1.0),
                ("e", mu[::2, 1], 0.7)]
    out, nodes = {}, []
    for nm, parent, prec in children:
        Y = GaussianARD(parent, prec, name="y_" + nm)
        y = rs.randn(*Y.plates)
        Y.observe(y)
        out["y_" + nm] = y
        out["plates_" + nm] = np.array(Y.plates)
        nodes.append(Y)
    Yv = GaussianARD(vec[1:4], [2.0, 0.5], shape=(2,), name="y_v")
    yv = rs.randn(*(Yv.plates + (2,)))
    Yv.observe(yv)
    out["y_v"] = yv
    Q = VB(mu, vec, Yv, *nodes)
    Q.update(repeat=2, verbose=False, tol=0)
    out["L"] = Q.L[:2].copy()
    node_state("mu", mu, out)
    node_state("vec", vec, out)
    save(name, **out)


def lssm_plated_dynamics(name="lssm_plated_dynamics", M=4, N=15, D=2, P=3, iters=4):
    """P independent chains, each with ITS OWN time-invariant dynamics: A with plates (P, 1, D)."""
    from bayespy.nodes import GaussianMarkovChain, Dot
    rs = np.random.RandomState(22)
    y = rs.randn(M, P, N).cumsum(axis=-1) * 0.3 + rs.randn(M, P, N)
    alpha = Gamma(1e-5, 1e-5, plates=(D,), name="alpha")
    A = GaussianARD(0, alpha, shape=(D,), plates=(P, 1, D), name="A")
    A_init = 0.5 * rs.randn(P, 1, D, D)
    A.initialize_from_value(A_init)
    X = GaussianMarkovChain(np.zeros(D), 1e-3 * np.identity(D), A, np.ones(D), n=N, name="X")
    assert X.plates == (P,)
    gamma = Gamma(1e-5, 1e-5, plates=(D,), name="gamma")
    C = GaussianARD(0, gamma, shape=(D,), plates=(M, 1, 1), name="C")
    F = Dot(C, X, name="F")
    C_init = rs.randn(M, 1, 1, D)
    C.initialize_from_value(C_init)
    tau = Gamma(1e-5, 1e-5, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    Y.observe(y)
    Q = VB(X, C, gamma, A, alpha, tau, Y)
    Q.update(repeat=iters, verbose=False, tol=0)
    out = dict(y=y, C_init=C_init, A_init=A_init, L=Q.L[:iters])
    for nm, node in (("X", X), ("C", C), ("A", A), ("alpha", alpha), ("tau", tau)):
        node_state(nm, node, out)
    save(name, **out)


def pca_bench_prefix(name="pca_bench_100k", N=100_000, iters=5):
    """The benchmark's own model and data (bench.py: synth_shard / init_C, M=64, K=16) on the first N columns:
    the reference trajectory of the data the GPU numbers are measured on.  N = 1e5 makes the multi-tile hand-off
    of the fused sweep kernel live (ntiles >= 2 x 148).  y is regenerated from the seed by the test."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    import bench
    M, K = bench.M_DIM, bench.K_DIM
    y = bench.synth_shard(M, 0, N, 1)
    X = GaussianARD(0, 1, plates=(1, N), shape=(K,))
    alpha = Gamma(1e-5, 1e-5, plates=(K,))
    C = GaussianARD(0, alpha, plates=(M, 1), shape=(K,))
    F = SumMultiply('d,d->', X, C)
    tau = Gamma(1e-5, 1e-5)
    Y = GaussianARD(F, tau)
    Y.observe(y)
    C.initialize_from_value(bench.init_C(M, K))
    Q = VB(Y, X, C, alpha, tau)
    Q.update(repeat=iters, verbose=False, tol=0)
    sel = np.r_[0:64, N // 2:N // 2 + 64, N - 64:N]
    out = dict(N=N, L=Q.L[:iters], X_sel=sel, X_u0_sel=np.asarray(X.u[0])[0, sel, :],
               X_cov=np.asarray(X.u[1])[0, 0] - np.outer(np.asarray(X.u[0])[0, 0], np.asarray(X.u[0])[0, 0]),
               y_checksum=np.array([y.sum(), (y * y).sum(), y[:, -1].sum()]))
    for k, node in (("Y", Y), ("X", X), ("C", C), ("alpha", alpha), ("tau", tau)):
        out["l_" + k] = Q.l[node][:iters]
    for nm, node in (("C", C), ("alpha", alpha), ("tau", tau)):
        node_state(nm, node, out)
    save(name, **out)


def take_models(name="take"):
    """nodes/take.py: group means picked by integer labels (messages are scatter-added back), a masked observation,
    and a vector-valued node taken along plate axis -2 with a matrix of indices."""
    from bayespy.nodes import Take
    rs = np.random.RandomState(31)
    N, G = 40, 3
    idx = rs.randint(0, G, size=N)
    idx[:3] = [-1, -3, 2]                                  # negative indices count from the end (np.take)
    y = np.array([-2.0, 0.5, 3.0])[idx % G] + 0.3 * rs.randn(N)
    mask = rs.rand(N) < 0.85
    mu = GaussianARD(0, 1e-3, plates=(G,), name="mu")
    m = Take(mu, idx, name="m")
    tau = Gamma(1e-3, 1e-3, name="tau")
    Y = GaussianARD(m, tau, name="Y")
    Y.observe(y, mask=mask)
    Q = VB(mu, tau, Y)
    Q.update(repeat=5, verbose=False, tol=0)
    out = dict(idx=idx, y=y, mask=mask, L=Q.L[:5], m_u0=np.asarray(m.get_moments()[0]), m_u1=np.asarray(m.get_moments()[1]))
    node_state("mu", mu, out)
    node_state("tau", tau, out)
    # vector-valued, plate axis -2, matrix of indices
    X = GaussianARD(0, 1, plates=(3, 4), shape=(2,), name="X")
    X.initialize_from_value(rs.randn(3, 4, 2))
    idx2 = np.array([[2, 0], [1, 1]])
    Z = Take(X, idx2, plate_axis=-2, name="Z")
    assert Z.plates == (2, 2, 4)
    y2 = rs.randn(2, 2, 4, 2)
    W = GaussianARD(Z, 2.0, name="W")
    W.observe(y2)
    Q2 = VB(X, W)
    Q2.update(repeat=2, verbose=False, tol=0)
    out.update(idx2=idx2, y2=y2, X_init=np.asarray(X.u[0]) * 0 + 0, L2=Q2.L[:2], Z_u0=np.asarray(Z.get_moments()[0]))
    node_state("X", X, out)
    save(name, **out)


def gate_models(name="gate"):
    """nodes/gate.py: a mixture written with Gate (scalar and vector-valued gated nodes, gated plate -1 and -2)."""
    from bayespy.nodes import Gate
    rs = np.random.RandomState(41)
    N, K = 50, 3
    z_true = rs.randint(0, K, size=N)
    y = np.array([-3.0, 0.0, 4.0])[z_true] + 0.5 * rs.randn(N)
    alpha = Dirichlet(np.ones(K), name="alpha")
    Z = Categorical(alpha, plates=(N,), name="Z")
    z_init = rs.randint(0, K, size=N)
    Z.initialize_from_value(z_init)
    mu = GaussianARD(0, 1e-3, plates=(K,), name="mu")
    mu.initialize_from_value(np.array([-1.0, 0.5, 2.0]))
    F = Gate(Z, mu, name="F")
    assert F.plates == (N,)
    tau = Gamma(1e-3, 1e-3, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    Y.observe(y)
    Q = VB(Z, mu, alpha, tau, Y)
    Q.update(repeat=5, verbose=False, tol=0)
    out = dict(y=y, z_init=z_init, L=Q.L[:5], F_u0=np.asarray(F.get_moments()[0]), F_u1=np.asarray(F.get_moments()[1]))
    for nm, node in (("Z", Z), ("mu", mu), ("alpha", alpha), ("tau", tau)):
        node_state(nm, node, out)
    # vector-valued node gated over plate axis -2
    X = GaussianARD(0, 1, plates=(K, 4), shape=(2,), name="X")
    X_init = rs.randn(K, 4, 2)
    X.initialize_from_value(X_init)
    Z2 = Categorical(np.ones(K) / K, plates=(6, 1), name="Z2")
    z2_init = rs.randint(0, K, size=(6, 1))
    Z2.initialize_from_value(z2_init)
    G2 = Gate(Z2, X, gated_plate=-2, name="G2")
    assert G2.plates == (6, 4)
    y2 = rs.randn(6, 4, 2)
    W = GaussianARD(G2, 1.5, name="W")
    W.observe(y2)
    Q2 = VB(Z2, X, W)
    Q2.update(repeat=3, verbose=False, tol=0)
    out.update(X_init=X_init, z2_init=z2_init, y2=y2, L2=Q2.L[:3], G2_u0=np.asarray(G2.get_moments()[0]))
    node_state("Z2", Z2, out)
    node_state("X", X, out)
    save(name, **out)


def lssm_doc_rotated(name="lssm_doc_rotated"):
    """doc/source/examples/lssm.rst in full: 10 plain iterations, then the rotation parameter expansion
    (transformations.py:1112-1452 RotateGaussianMarkovChain with a plate-rotating RotateGaussianARD for the dynamics,
    :376-1110) as callback until convergence.  Pins the cost functions and their gradients (R and Q) from a known
    state as well."""
    from bayespy.nodes import GaussianMarkovChain, Dot
    from bayespy.inference.vmp import transformations
    np.random.seed(1)
    M, N, D = 30, 400, 10
    alpha = Gamma(1e-5, 1e-5, plates=(D,), name="alpha")
    A = GaussianARD(0, alpha, shape=(D,), plates=(D,), name="A")
    X = GaussianMarkovChain(np.zeros(D), 1e-3 * np.identity(D), A, np.ones(D), n=N, name="X")
    gamma = Gamma(1e-5, 1e-5, plates=(D,), name="gamma")
    C = GaussianARD(0, gamma, shape=(D,), plates=(M, 1), name="C")
    F = Dot(C, X, name="F")
    C.initialize_from_random()
    C_init = np.asarray(C.u[0]).copy()
    tau = Gamma(1e-5, 1e-5, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    Q = VB(X, C, gamma, A, alpha, tau, Y)
    w = 0.3
    a = np.array([[np.cos(w), -np.sin(w), 0, 0], [np.sin(w), np.cos(w), 0, 0], [0, 0, 1, 0], [0, 0, 0, 0]])
    c = np.random.randn(M, 4)
    x = np.empty((N, 4)); f = np.empty((M, N)); y = np.empty((M, N))
    x[0] = 10 * np.random.randn(4)
    f[:, 0] = np.dot(c, x[0])
    y[:, 0] = f[:, 0] + 3 * np.random.randn(M)
    for n in range(N - 1):
        x[n + 1] = np.dot(a, x[n]) + [1, 1, 10, 10] * np.random.randn(4)
        f[:, n + 1] = np.dot(c, x[n + 1])
        y[:, n + 1] = f[:, n + 1] + 3 * np.random.randn(M)
    mask = random.mask(M, N, p=0.2)
    Y.observe(y, mask=mask)
    Q.update(repeat=10, verbose=False)
    rotC = transformations.RotateGaussianARD(C, gamma)
    rotA = transformations.RotateGaussianARD(A, alpha)
    rotX = transformations.RotateGaussianMarkovChain(X, rotA)
    R = transformations.RotationOptimizer(rotX, rotC, D)
    # cost functions and gradients from the state after 10 iterations, at a fixed test rotation
    Rt = np.identity(D) + 0.1 * np.random.RandomState(3).randn(D, D)
    inv, logdet = np.linalg.inv(Rt), np.linalg.slogdet(Rt)[1]
    rotX.setup(); rotC.setup()
    bX, dbX = rotX.bound(Rt, inv=inv, logdet=logdet)
    bXonly, dbXonly = rotX._compute_bound(Rt, logdet=logdet, inv=inv, gradient=True)
    bA, dRA, dQA = rotA.bound(inv.T, inv=Rt.T, logdet=-logdet, Q=Rt)
    bC, dbC = rotC.bound(inv.T, inv=Rt.T, logdet=-logdet)
    out = dict(y=y, mask=np.asarray(mask), C_init=C_init, Rt=Rt, bX=np.asarray(bX), dbX=dbX, bXonly=np.asarray(bXonly),
               dbXonly=dbXonly, bA=np.asarray(bA), dRA=dRA, dQA=dQA, bC=np.asarray(bC), dbC=dbC, L10=Q.L[:10].copy())
    # one explicit rotation by Rt: the rotated state
    rotX.rotate(Rt, inv=inv, logdet=logdet)
    rotC.rotate(inv.T, inv=Rt.T, logdet=-logdet)
    for nm, node in (("X", X), ("A", A), ("alpha", alpha), ("C", C), ("gamma", gamma)):
        tmp = {}
        node_state(nm, node, tmp)
        for k, v in tmp.items():
            if nm == "X" and np.ndim(v) >= 2 and np.shape(v)[0] >= N - 1:
                v = v[::37]                     # a sample of the time steps keeps the fixture small
            out["rot1_" + k] = np.array(v, copy=True)      # the reference updates node.u in place later on
    out["rot1_bound"] = Q.compute_lowerbound()
    Q.callback = R.rotate
    Q.update(repeat=1000, verbose=False)
    out["L"] = Q.L[:Q.iter].copy()
    out["iters"] = Q.iter
    for nm, node in (("A", A), ("alpha", alpha), ("C", C), ("gamma", gamma), ("tau", tau)):
        node_state(nm, node, out)
    out["X_u0_sel"] = np.asarray(X.u[0])[::37]
    save(name, **out)


def gaussian_gamma_models(name="gaussian_gamma"):
    """nodes/gaussian.py GaussianGamma: (a) a factor model whose loadings carry a per-row gamma scale, through
    SumMultiply (Gaussian-gamma output) into an observed GaussianARD; (b) a Gaussian-gamma mean with unknown
    mean / precision / rate under a Gaussian likelihood; (c) scalar (ndim=0) Gaussian-gamma variables; (d) a diagonal
    Wishart made of gamma scalars; (e) rotate / translate of q."""
    from bayespy.nodes import GaussianGamma
    rs = np.random.RandomState(11)
    out = {}
    # (a)
    M, N, K = 5, 30, 3
    y = rs.randn(M, K) @ rs.randn(K, N) + 0.2 * rs.randn(M, N)
    mask = rs.rand(M, N) < 0.9
    b = Gamma(2.0, 2.0, plates=(M, 1), name="b")
    W = GaussianGamma(np.zeros(K), np.identity(K), 3.0, b, plates=(M, 1), name="W")
    X = GaussianARD(0, 1, shape=(K,), plates=(1, N), name="X")
    Xi = rs.randn(1, N, K)
    X.initialize_from_value(Xi)
    F = SumMultiply("k,k->", W, X, name="F")
    assert F.dims == ((), (), (), ())
    tau = Gamma(1e-3, 1e-3, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    Y.observe(y, mask=mask)
    Q = VB(Y, W, X, tau, b)
    Q.update(repeat=6, verbose=False, tol=0)
    out.update(a_y=y, a_mask=mask, a_Xinit=Xi, a_L=Q.L[:6].copy())
    for nm, nd in (("a_W", W), ("a_X", X), ("a_tau", tau), ("a_b", b)):
        node_state(nm, nd, out)
    for i, u in enumerate(F.get_moments()):
        out["a_F_u%d" % i] = np.array(u, copy=True)
    for nm, nd in (("a_lW", W), ("a_lX", X), ("a_ltau", tau), ("a_lb", b), ("a_lY", Y)):
        out[nm] = nd.lower_bound_contribution()
    # (b)
    Dm, Nb = 3, 25
    z = rs.randn(Nb, Dm) + np.array([1.0, -2.0, 0.5])
    mu0 = Gaussian(np.zeros(Dm), 1e-2 * np.identity(Dm), name="mu0")
    L0 = Wishart(Dm + 1.0, np.identity(Dm), name="L0")
    bb = Gamma(1.5, 1.0, name="bb")
    m = GaussianGamma(mu0, L0, 2.5, bb, name="m")
    L1 = Wishart(Dm + 2.0, np.identity(Dm), name="L1")
    Z = Gaussian(m, L1, plates=(Nb,), name="Z")
    Z.observe(z)
    Q = VB(Z, m, L1, mu0, L0, bb)
    Q.update(repeat=5, verbose=False, tol=0)
    out.update(b_z=z, b_L=Q.L[:5].copy())
    for nm, nd in (("b_m", m), ("b_L1", L1), ("b_mu0", mu0), ("b_L0", L0), ("b_bb", bb)):
        node_state(nm, nd, out)
    # (c) scalars: GaussianGamma(ndim=0) as the mean of a GaussianARD
    P = 4
    yc = rs.randn(6, P) * 0.5 + np.arange(P)
    lam = Gamma(2.0, 1.0, plates=(P,), name="lam")
    bc = Gamma(1.0, 1.0, plates=(P,), name="bc")
    mc = GaussianGamma(np.zeros(P), lam.as_wishart(ndim=0), 1.5 * np.ones(P), bc, ndim=0, name="mc")
    assert mc.plates == (P,) and mc.dims == ((), (), (), ())
    al = Gamma(1e-2, 1e-2, plates=(6, 1), name="al")
    Yc = GaussianARD(mc, al, name="Yc")
    assert Yc.plates == (6, P)
    Yc.observe(yc)
    Q = VB(Yc, mc, lam, bc, al)
    Q.update(repeat=5, verbose=False, tol=0)
    out.update(c_y=yc, c_L=Q.L[:5].copy())
    for nm, nd in (("c_mc", mc), ("c_lam", lam), ("c_bc", bc), ("c_al", al)):
        node_state(nm, nd, out)
    # (d) diagonal Wishart from gamma scalars as the precision of a Gaussian
    g = Gamma(1e-2, 1e-2, plates=(Dm,), name="g")
    Zd = Gaussian(np.zeros(Dm), g.diag(), plates=(Nb,), name="Zd")
    Zd.observe(z)
    Q = VB(Zd, g)
    Q.update(repeat=2, verbose=False, tol=0)
    out.update(d_L=Q.L[:2].copy())
    node_state("d_g", g, out)
    for i, u in enumerate(g.diag().get_moments()):
        out["d_W_u%d" % i] = np.array(u, copy=True)
    # (e) rotate and translate a Gaussian-gamma posterior
    R = rs.randn(Dm, Dm)
    bvec = rs.randn(Dm)
    m.rotate(R)
    node_state("e_rot", m, out)
    m.translate(bvec)
    node_state("e_tra", m, out)
    out.update(e_R=R, e_b=bvec, e_loc=m.get_gaussian_location())
    # (get_gaussian_mean_and_variance raises AttributeError in the reference: it reads self.ndim, which the node lacks)
    save(name, **out)


def lssm_inputs(name="lssm_inputs", M=4, N=25, D=2, K=2, P=3, iters=5):
    """gaussian_markov_chain.py:485-540, :608-616, :638-655: a state-space model driven by input signals.  (a) known
    inputs (an array), dynamics [A B] learned; (b) uncertain inputs (a Gaussian node that receives messages), per-chain
    plates, a masked observation; (c) dynamics with a gamma scale of their own (a GaussianGamma node as A)."""
    from bayespy.nodes import GaussianMarkovChain, GaussianGamma
    rs = np.random.RandomState(17)
    out = {}
    # (a)
    z = np.stack([np.sin(0.4 * np.arange(N - 1)), np.ones(N - 1)], axis=-1)           # (N-1, K)
    x = np.zeros((N, D))
    a_true = np.array([[0.8, -0.3], [0.3, 0.8]])
    b_true = np.array([[1.0, 0.2], [0.0, -0.5]])
    for n in range(N - 1):
        x[n + 1] = a_true @ x[n] + b_true @ z[n] + 0.1 * rs.randn(D)
    c = rs.randn(M, D)
    y = c @ x.T + 0.2 * rs.randn(M, N)
    alpha = Gamma(1e-3, 1e-3, plates=(D + K,), name="alpha")
    A = GaussianARD(0, alpha, shape=(D + K,), plates=(D,), name="A")
    X = GaussianMarkovChain(np.zeros(D), 1e-2 * np.identity(D), A, np.ones(D), inputs=z, n=N, name="X")
    C = GaussianARD(0, 1e-2, shape=(D,), plates=(M, 1), name="C")
    C_init = rs.randn(M, 1, D)
    C.initialize_from_value(C_init)
    F = SumMultiply("i,i", C, X, name="F")
    tau = Gamma(1e-3, 1e-3, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    Y.observe(y)
    Q = VB(X, C, A, alpha, tau, Y)
    Q.update(repeat=iters, verbose=False, tol=0)
    out.update(a_z=z, a_y=y, a_Cinit=C_init, a_L=Q.L[:iters].copy())
    for nm, nd in (("a_X", X), ("a_C", C), ("a_A", A), ("a_alpha", alpha), ("a_tau", tau)):
        node_state(nm, nd, out)
    # (b) uncertain inputs, chain plates
    U_mean = rs.randn(P, N - 1, K)
    U = GaussianARD(U_mean, 4.0, shape=(K,), plates=(P, N - 1), name="U")
    A2 = GaussianARD(0, 1.0, shape=(D + K,), plates=(P, 1, D), name="A2")
    A2_init = 0.3 * rs.randn(P, 1, D, D + K)
    A2.initialize_from_value(A2_init)
    nu2 = Gamma(2.0, 2.0, plates=(P, 1, D), name="nu2")
    X2 = GaussianMarkovChain(np.zeros(D), np.identity(D), A2, nu2, inputs=U, name="X2")
    assert X2.plates == (P,)
    y2 = rs.randn(P, N, D)
    mask2 = rs.rand(P, N) < 0.8
    Y2 = Gaussian(X2, 5.0 * np.identity(D), name="Y2")
    Y2.observe(y2, mask=mask2)
    Q = VB(X2, A2, nu2, U, Y2)
    Q.update(repeat=iters, verbose=False, tol=0)
    out.update(b_Umean=U_mean, b_y=y2, b_mask=mask2, b_A2init=A2_init, b_L=Q.L[:iters].copy())
    for nm, nd in (("b_X2", X2), ("b_A2", A2), ("b_nu2", nu2), ("b_U", U)):
        node_state(nm, nd, out)
    # (c) Gaussian-gamma dynamics
    b3 = Gamma(2.0, 1.0, plates=(D,), name="b3")
    A3 = GaussianGamma(np.zeros(D + K), np.identity(D + K), 2.0, b3, plates=(D,), name="A3")
    X3 = GaussianMarkovChain(np.zeros(D), np.identity(D), A3, np.ones(D), inputs=z, n=N, name="X3")
    Y3 = Gaussian(X3, 10.0 * np.identity(D), name="Y3")
    y3 = x + 0.3 * rs.randn(N, D)
    Y3.observe(y3)
    Q = VB(X3, A3, b3, Y3)
    Q.update(repeat=iters, verbose=False, tol=0)
    out.update(c_y=y3, c_L=Q.L[:iters].copy())
    for nm, nd in (("c_X3", X3), ("c_A3", A3), ("c_b3", b3)):
        node_state(nm, nd, out)
    save(name, **out)


def lssm_mixing_plated(name="lssm_mixing_plated", P=3, N=10, D=2, K=2, iters=4):
    """VaryingGaussianMarkovChain over chain plates: mixing matrices with plates (P, D), weights with plates (P, N-1),
    an innovation precision with plates (P, 1, D); a masked Gaussian observation of the states."""
    from bayespy.nodes import VaryingGaussianMarkovChain
    rs = np.random.RandomState(44)
    B = GaussianARD(0, 0.5, shape=(D, K), plates=(P, D), name="B")
    B_init = 0.5 * rs.randn(P, D, D, K)
    B.initialize_from_value(B_init)
    S = GaussianARD(0, 1, shape=(K,), plates=(P, N - 1), name="S")
    S_init = rs.randn(P, N - 1, K)
    S.initialize_from_value(S_init)
    nu = 1.0 + rs.rand(P, 1, D)
    X = VaryingGaussianMarkovChain(np.zeros(D), np.identity(D), B, S, nu, name="X")
    assert X.plates == (P,) and X.dims[0] == (N, D)
    y = rs.randn(P, N, D)
    mask = rs.rand(P, N) < 0.8
    Y = Gaussian(X, 4.0 * np.identity(D), name="Y")
    Y.observe(y, mask=mask)
    Q = VB(X, B, S, Y)
    Q.update(repeat=iters, verbose=False, tol=0)
    out = dict(y=y, mask=mask, B_init=B_init, S_init=S_init, nu=nu, L=Q.L[:iters].copy())
    for nm, node in (("X", X), ("B", B), ("S", S)):
        node_state(nm, node, out)
    save(name, **out)


def lssm_varying_rotated(name="lssm_varying_rotated", M=4, N=16, D=2, K=2, P=2, iters=4):
    """transformations.py:1454-1541 with the array form of RotateGaussianARD (:376-1110): (a) a chain with mixed dynamics
    A_n = sum_k s_nk B_k rotated together with its loadings after every VB iteration (RotateVaryingMarkovChain, the
    mixing matrices rotated on variable axis -2 and plate axis -1 with an ARD precision over both variable axes);
    (b) plated chains with per-chain time-varying dynamics (RotateGaussianMarkovChain)."""
    from bayespy.nodes import GaussianMarkovChain
    from bayespy.inference.vmp.transformations import (RotateGaussianARD, RotateVaryingMarkovChain,
                                                       RotateGaussianMarkovChain, RotationOptimizer)
    rs = np.random.RandomState(55)
    out = {}
    y = rs.randn(M, N).cumsum(axis=-1) * 0.4 + 0.3 * rs.randn(M, N)
    beta = Gamma(1e-3, 1e-3, plates=(D, K), name="beta")
    B = GaussianARD(0, beta, shape=(D, K), plates=(1, D), name="B")
    B_init = 0.5 * rs.randn(1, D, D, K)
    B.initialize_from_value(B_init)
    S = GaussianARD(0, 1, shape=(K,), plates=(N - 1, 1), name="S")
    S_init = rs.randn(N - 1, 1, K)
    S.initialize_from_value(S_init)
    A = SumMultiply("dk,k->d", B, S, name="A")
    X = GaussianMarkovChain(np.zeros(D), 1e-3 * np.identity(D), A, np.ones(D), n=N, name="X")
    gamma = Gamma(1e-3, 1e-3, plates=(D,), name="gamma")
    C = GaussianARD(0, gamma, shape=(D,), plates=(M, 1), name="C")
    C_init = rs.randn(M, 1, D)
    C.initialize_from_value(C_init)
    F = SumMultiply("d,d", C, X, name="F")
    tau = Gamma(1e-3, 1e-3, name="tau")
    Y = GaussianARD(F, tau, name="Y")
    Y.observe(y)
    Q = VB(Y, X, C, gamma, B, beta, S, tau)
    rotB = RotateGaussianARD(B, beta, axis=-2)
    rotX = RotateVaryingMarkovChain(X, B, S, rotB)
    rotC = RotateGaussianARD(C, gamma)
    Rot = RotationOptimizer(rotX, rotC, D)
    Ls = []
    for it in range(iters):
        Q.update(verbose=False)
        Rot.rotate(maxiter=10)
        Ls.append(Q.compute_lowerbound())
    out.update(a_y=y, a_Binit=B_init, a_Sinit=S_init, a_Cinit=C_init, a_L=Q.L[:iters].copy(), a_Lrot=np.array(Ls))
    for nm, nd in (("a_X", X), ("a_C", C), ("a_B", B), ("a_beta", beta), ("a_gamma", gamma), ("a_tau", tau)):
        node_state(nm, nd, out)
    # (b) plated chains, per-chain time-varying dynamics; an observation per chain
    alpha = Gamma(1e-3, 1e-3, plates=(D,), name="alpha")
    A2 = GaussianARD(0, alpha, shape=(D,), plates=(P, N - 1, D), name="A2")
    A2_init = 0.4 * rs.randn(P, N - 1, D, D)
    A2.initialize_from_value(A2_init)
    X2 = GaussianMarkovChain(np.zeros(D), 1e-2 * np.identity(D), A2, np.ones(D), name="X2")
    assert X2.plates == (P,)
    gamma2 = Gamma(1e-3, 1e-3, plates=(D,), name="gamma2")
    C2 = GaussianARD(0, gamma2, shape=(D,), plates=(M, 1, 1), name="C2")
    C2_init = rs.randn(M, 1, 1, D)
    C2.initialize_from_value(C2_init)
    F2 = SumMultiply("d,d", C2, X2, name="F2")
    y2 = rs.randn(M, P, N)
    Y2 = GaussianARD(F2, 3.0, name="Y2")
    Y2.observe(y2)
    Q2 = VB(Y2, X2, C2, gamma2, A2, alpha)
    rotA2 = RotateGaussianARD(A2, alpha)
    rotX2 = RotateGaussianMarkovChain(X2, rotA2)
    rotC2 = RotateGaussianARD(C2, gamma2)
    Rot2 = RotationOptimizer(rotX2, rotC2, D)
    Ls2 = []
    for it in range(iters):
        Q2.update(verbose=False)
        Rot2.rotate(maxiter=10)
        Ls2.append(Q2.compute_lowerbound())
    out.update(b_y=y2, b_A2init=A2_init, b_C2init=C2_init, b_L=Q2.L[:iters].copy(), b_Lrot=np.array(Ls2))
    for nm, nd in (("b_X2", X2), ("b_C2", C2), ("b_A2", A2), ("b_alpha", alpha), ("b_gamma2", gamma2)):
        node_state(nm, nd, out)
    save(name, **out)


def multinomial_models(name="multinomial"):
    """nodes/multinomial.py: (a) counts with a Dirichlet prior (posterior in closed form), (b) a mixture of multinomials
    with a different number of trials per item (the trials array has a unit axis where the cluster axis is)."""
    from bayespy.nodes import Multinomial
    rs = np.random.RandomState(9)
    K, C, N = 4, 3, 40
    out = {}
    p = Dirichlet(np.array([1.0, 2.0, 0.5, 1.5]), name="p")
    x = rs.multinomial(12, [0.1, 0.4, 0.2, 0.3], size=6)
    X = Multinomial(12, p, plates=(6,), name="X")
    X.observe(x)
    Q = VB(X, p)
    Q.update(repeat=2, verbose=False, tol=0)
    out.update(a_x=x, a_L=Q.L[:2].copy())
    node_state("a_p", p, out)
    # (b)
    ptrue = rs.dirichlet(np.ones(K), size=C)
    lab = rs.randint(0, C, size=N)
    n = rs.randint(5, 30, size=(N, 1))
    counts = np.array([rs.multinomial(n[i, 0], ptrue[lab[i]]) for i in range(N)])
    alpha = Dirichlet(np.ones(C), name="alpha")
    Z = Categorical(alpha, plates=(N,), name="Z")
    z_init = rs.randint(0, C, size=N)
    Z.initialize_from_value(z_init)
    P = Dirichlet(np.ones(K), plates=(C,), name="P")
    Xm = Mixture(Z, Multinomial, n, P, name="Xm")
    Xm.observe(counts)
    Q = VB(Xm, P, Z, alpha)
    Q.update(repeat=6, verbose=False, tol=0)
    out.update(b_n=n, b_counts=counts, b_zinit=z_init, b_L=Q.L[:6].copy())
    for nm, nd in (("b_P", P), ("b_Z", Z), ("b_alpha", alpha)):
        node_state(nm, nd, out)
    # an unobserved multinomial: moments and log-normaliser
    Xf = Multinomial(np.array([[3], [7]]), np.array([[0.2, 0.8], [0.5, 0.5], [0.9, 0.1]]), name="Xf")
    assert Xf.plates == (2, 3)
    node_state("c_Xf", Xf, out)
    save(name, **out)


def reference_demos(name="demos"):
    """bayespy/demos/*.py run as they are (matplotlib answered by a no-op stand-in): every bound the scripts print."""
    import contextlib
    import importlib
    import io
    import re
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from bayespy_b200 import _Permissive
    for mod in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.colors", "matplotlib.patches",
                "matplotlib.gridspec"):
        sys.modules.setdefault(mod, _Permissive())
    from demo_calls import CALLS
    out = {}
    for demo, call in CALLS.items():
        m = importlib.import_module("bayespy.demos." + demo)
        buf = io.StringIO()
        np.random.seed(1)
        try:
            with contextlib.redirect_stdout(buf):
                call(m)
        except RuntimeError as err:           # Q.save() in a container without HDF5: no reference output for this demo
            print(demo, "not recorded:", err)
            continue
        L = [float(v) for v in re.findall(r"(?:loglike=|integrated pdf: )([-+]?(?:[0-9.]+(?:e[-+][0-9]+)?|inf|nan))", buf.getvalue())]
        print(demo, len(L), L[:2], L[-1:])
        out[demo] = np.array(L)
    save(name, **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["quickstart", "pca", "linalg", "summul", "dist", "gmm", "gmc", "rot", "dot", "mixard", "gmcplates", "gmcvarying", "pcabench", "pcamasked64", "take", "gate", "lssmrot", "gmcmixing", "gradients", "lda", "slice", "gg", "gmcinputs", "rotgeneral", "multinomial", "demos"]
    if "quickstart" in which:
        quickstart()
    if "pca" in which:
        pca("pca_small", 20, 100, 5)
        pca("pca_64x16", 64, 96, 16, iters=5)
        pca("pca_masked", 12, 40, 4, mask_p=0.8)
    if "linalg" in which:
        linalg_vectors()
    if "summul" in which:
        summul_vectors()
    if "dist" in which:
        distribution_vectors()
    if "gmm" in which:
        gmm("gmm_small", 300, 3, 5)
        gmm_doc()
    if "mixard" in which:
        mixture_ard()
    if "dot" in which:
        summultiply_nodes()
    if "rot" in which:
        pca_rotated()
    if "gmcplates" in which:
        lssm_plated()
    if "gate" in which:
        gate_models()
    if "lssmrot" in which:
        lssm_doc_rotated()
    if "slice" in which:
        sliced_nodes()
    if "lda" in which:
        lda()
        advanced_guide()
    if "gradients" in which:
        pca_gradients()
        svi_mixture()
    if "gmcmixing" in which:
        lssm_mixing_plated()
        lssm_mixing()
        lssm_switching()
    if "take" in which:
        take_models()
    if "gg" in which:
        gaussian_gamma_models()
    if "gmcinputs" in which:
        lssm_inputs()
    if "rotgeneral" in which:
        lssm_varying_rotated()
    if "multinomial" in which:
        multinomial_models()
    if "demos" in which:
        reference_demos()
    if "pcamasked64" in which:
        pca("pca_masked_64x16", 64, 300, 16, mask_p=0.8, iters=4)
    if "pcabench" in which:
        pca_bench_prefix()
    if "gmcvarying" in which:
        lssm_varying()
        lssm_plated_dynamics()
    if "gmc" in which:
        block_banded_vectors()
        lssm("lssm_small", 6, 40, 3)
        lssm("lssm_masked", 6, 40, 3, mask_p=0.7)
