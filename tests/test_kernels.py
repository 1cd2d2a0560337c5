"""Per-kernel parity.  With backend='oracle' (CPU) these pin the oracle against the
reference's golden vectors (tests/golden/make_golden.py); with backend='cuda' (-m gpu)
they are the parity tests of the CUDA kernels, through the same C-ABI wrappers.
Tolerance: 1e-9 relative for fp64 kernels (north_star requires 1e-5)."""
import numpy as np
import pytest

from conftest import golden
from bayespy_b200.darray import DArray

RTOL = 1e-9


def close(a, b, rtol=RTOL, atol=1e-11):
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=rtol, atol=atol)


# ---- utils.linalg (linalg.py:31-223) ------------------------------------------------------------
@pytest.mark.parametrize("D", [1, 3, 8, 16, 33])
def test_linalg_golden(backend, D):
    from bayespy_b200.utils import linalg
    g = golden("linalg")
    A, b, B = g["A%d" % D], g["b%d" % D], g["B%d" % D]
    U = linalg.chol(A)
    close(np.triu(U), g["U%d" % D])
    close(linalg.chol_solve(U, b), g["solve%d" % D])
    close(linalg.chol_solve(U, B, matrix=True), g["solveM%d" % D])
    close(linalg.chol_inv(U), g["inv%d" % D])
    close(linalg.chol_logdet(U), g["logdet%d" % D])


def test_chol_not_spd_raises(backend):
    from bayespy_b200.utils import linalg
    A = np.identity(4)[None] * np.ones((3, 1, 1))
    A[1, 2, 2] = -1.0
    with pytest.raises(Exception, match="Matrix not positive definite"):
        linalg.chol(A)


def test_linalg_scalar_paths(backend):
    """ndim=0 fast paths (linalg.py:40-41,71-72,176-177,211-212)."""
    from bayespy_b200.utils import linalg
    c = np.array([4.0, 9.0, 2.5])
    U = linalg.chol(c, ndim=0)
    close(U, np.sqrt(c))
    close(linalg.chol_solve(U, np.array([1.0, 2.0, 3.0]), ndim=0), np.array([1.0, 2.0, 3.0]) / c)
    close(linalg.chol_inv(U, ndim=0), 1 / c)
    close(linalg.chol_logdet(U, ndim=0), np.log(c))


# ---- utils.misc (misc.py:805-945) ---------------------------------------------------------------
def test_sum_multiply_golden(backend):
    from bayespy_b200.utils import misc
    g = golden("summul")
    a, b, c, A = g["a"], g["b"], g["c"], g["A"]
    close(misc.sum_multiply(a, b, c), g["r0"])
    close(misc.sum_multiply(a, b, c, axis=-1), g["r1"])
    close(misc.sum_multiply(a, b, c, axis=(0, 2), keepdims=True), g["r2"])
    close(misc.sum_multiply(a, b, axis=[1], sumaxis=False), g["r3"])
    close(misc.sum_product(a, b, c, axes_to_keep=[0, 2]), g["r4"])
    close(misc.sum_product(a, b, axes_to_sum=[-2], keepdims=True), g["r5"])
    for k, r in (("p0", misc.sum_multiply_to_plates(a, b, to_plates=(3, 1), from_plates=(4, 3, 5))),
                 ("p1", misc.sum_multiply_to_plates(a, b, to_plates=(1,), from_plates=(4, 3, 5))),
                 ("p2", misc.sum_multiply_to_plates(c, to_plates=(), from_plates=(6, 5))),
                 ("p3", misc.sum_multiply_to_plates(A, to_plates=(3,), from_plates=(4, 3), ndim=2))):
        assert np.shape(r) == np.shape(g[k]), k
        close(r, g[k])


def test_sum_multiply_bruteforce(backend):
    """test_misc.py:157-340 pattern: against explicit product + sum for many axis choices."""
    from bayespy_b200.utils import misc
    rng = np.random.RandomState(5)
    x, y, z = rng.randn(2, 3, 4), rng.randn(3, 1), rng.randn(4)
    full = x * y * z
    for axis in (None, 0, 1, 2, -1, (0, 1), (0, 2), (1, 2), (0, 1, 2)):
        for keep in (False, True):
            ref = np.sum(full, axis=axis, keepdims=keep)
            close(misc.sum_multiply(x, y, z, axis=axis, keepdims=keep), ref)
    with pytest.raises(ValueError):
        misc.sum_multiply()
    with pytest.raises(ValueError):
        misc.sum_multiply(x, axis=[5], sumaxis=False)


def test_sum_multiply_large_reduction(backend):
    """Exercises the split (block + final) reduction path."""
    from bayespy_b200.utils import misc
    rng = np.random.RandomState(6)
    a = rng.randn(3, 40000)
    b = rng.randn(40000)
    close(misc.sum_multiply(a, b, axis=-1), (a * b).sum(-1), rtol=1e-10)
    m = rng.rand(3, 40000) > 0.3
    from bayespy_b200 import darray as D
    r = D.reduce_to_shape(D.asarray(a), (3, 1), mask=D.DArray.from_numpy(m), from_shape=(3, 40000))
    close(r.numpy(), (a * m).sum(-1, keepdims=True), rtol=1e-10)


def test_ewise_ops(backend):
    from bayespy_b200 import darray as D
    import scipy.special as sp
    rng = np.random.RandomState(8)
    a = rng.gamma(2.0, 2.0, size=(5, 1, 7)) + 0.05
    b = rng.randn(4, 7)
    A, B = D.asarray(a), D.asarray(b)
    close((A + B).numpy(), a + b)
    close((A - B).numpy(), a - b)
    close((A * B).numpy(), a * b)
    close((B / A).numpy(), b / a)
    close((2.0 - A).numpy(), 2 - a)
    close((1.0 / A).numpy(), 1 / a)
    close(D.axpby(2.0, A, -3.0, B).numpy(), 2 * a - 3 * b)
    close(D.fma(0.5, A, B, 2.0, A).numpy(), 0.5 * a * b + 2 * a)
    close(D.log(A).numpy(), np.log(a))
    close(D.exp(B).numpy(), np.exp(b))
    close(D.sqrt(A).numpy(), np.sqrt(a))
    close(D.gammaln(A).numpy(), sp.gammaln(a), rtol=1e-12, atol=1e-13)
    close(D.digamma(A).numpy(), sp.digamma(a), rtol=1e-11, atol=1e-13)
    big = np.array([1e-3, 0.3, 1.4616321449683623, 7.0, 55.5, 1e4, 1e8])
    close(D.digamma(D.asarray(big)).numpy(), sp.digamma(big), rtol=1e-11, atol=1e-12)
    close(D.multigammaln(D.asarray(a + 3), 4).numpy(), np.vectorize(lambda v: sp.multigammaln(v, 4))(a + 3), rtol=1e-12)
    close(D.multidigamma(D.asarray(a + 3), 4).numpy(),
          np.sum(sp.digamma((a + 3)[..., None] - 0.5 * np.arange(4)), axis=-1), rtol=1e-11)
    m = rng.rand(5, 1, 7) > 0.5
    close(D.where(D.DArray.from_numpy(m), A, B).numpy(), np.where(m, a, b))
    # strided views: diagonal write, transpose, broadcast materialisation
    k = rng.randn(2, 3, 3)
    Kd = D.asarray(k)
    close(Kd.diag_view().numpy(), np.einsum("nii->ni", k))
    close(Kd.swap_last2().numpy(), np.swapaxes(k, -1, -2))
    Z = D.DArray.zeros((2, 3, 3))
    D.copy_into(Z.diag_view(), D.asarray(np.arange(6.0).reshape(2, 3)))
    close(Z.numpy(), np.einsum("ni,ij->nij", np.arange(6.0).reshape(2, 3), np.identity(3)))
    close(D.asarray(b).broadcast_to((3, 4, 7)).contiguous().numpy(), np.broadcast_to(b, (3, 4, 7)))


# ---- Distribution.compute_moments_and_cgf golden vectors ---------------------------------------
def test_gamma_moments_golden(backend):
    from bayespy_b200 import darray as D
    from bayespy_b200.engine.gamma import GammaDistribution
    g = golden("distributions")
    u, cgf = GammaDistribution().compute_moments_and_cgf([D.asarray(g["gam_phi0"]), D.asarray(g["gam_phi1"])])
    close(u[0].numpy(), g["gam_u0"]); close(u[1].numpy(), g["gam_u1"]); close(cgf.numpy(), g["gam_g"])

@pytest.mark.parametrize("shape_a,shape_b", [((6, 8), (6, 8)), ((5, 4, 4), (1, 4, 4)), ((5, 4, 4), (5, 1, 1)), ((64,), ()),
                                            ((7, 10), (10,)), ((7, 10), (7, 1)), ((3, 5, 6), (3, 1, 6)), ((9, 7), (9, 7)),
                                            ((2, 3, 4, 6), (1, 3, 1, 6)), ((1000, 32, 32), (32, 32))])
def test_ewise_fast_path_broadcasts(backend, shape_a, shape_b):
    """Even inner extents with contiguous / broadcast operands take the 16-byte path of bpk_ewise (one outer axis at most
    after collapsing); everything else the generic kernel.  Same results either way, bit for bit (pure elementwise)."""
    from bayespy_b200 import darray as D
    rng = np.random.RandomState(3)
    a, b = rng.randn(*shape_a), rng.randn(*shape_b) + 3.0
    A, B = D.asarray(a), D.asarray(b)
    np.testing.assert_array_equal((A + B).numpy(), a + b)
    np.testing.assert_array_equal((A * B).numpy(), a * b)
    np.testing.assert_array_equal((B - A).numpy(), b - a)
    np.testing.assert_array_equal((A / B).numpy(), a / b)
    close(D.axpby(2.0, A, -3.0, B).numpy(), 2.0 * a + -3.0 * b, rtol=1e-15, atol=1e-15)      # fused multiply-add on device
    close(D.affine(A, 1.5, 0.25).numpy(), 1.5 * a + 0.25, rtol=1e-15, atol=1e-15)
    got = D.fma(0.5, A, B, 2.0, A).numpy()
    close(got, 0.5 * a * b + 2.0 * a, rtol=1e-15, atol=1e-15)
    # a view that starts 8 bytes into its buffer is not 16-byte aligned: generic kernel
    if a.ndim == 1:
        v = D.asarray(np.concatenate([[0.0], a]))
        odd = v.slice_axis(0, 1, 1 + a.size)
        np.testing.assert_array_equal((odd + B).numpy(), a + b)



def test_gamma_domain_error(backend):
    from bayespy_b200 import darray as D
    from bayespy_b200.engine.gamma import GammaDistribution
    with pytest.raises((ValueError, FloatingPointError)):
        GammaDistribution().compute_moments_and_cgf([D.asarray([1.0, -1.0]), D.asarray([1.0, 1.0])])


def test_gaussian_moments_golden(backend):
    from bayespy_b200 import darray as D
    from bayespy_b200.engine.gaussian import GaussianARDDistribution
    g = golden("distributions")
    for pre in ("gau", "gaus"):      # per-plate covariance / shared covariance
        d = GaussianARDDistribution((5,))
        u, cgf = d.compute_moments_and_cgf([D.asarray(g[pre + "_phi0"]), D.asarray(g[pre + "_phi1"])])
        close(u[0].numpy(), g[pre + "_u0"])
        close(np.asarray(u[1]), g[pre + "_u1"])
        close(cgf.numpy(), g[pre + "_g"])


def test_raw_moment_kernels_golden(backend):
    """wishart / dirichlet / softmax entry points against the reference distributions."""
    from bayespy_b200.darray import DArray
    g = golden("distributions")
    be = backend
    # Wishart
    p0, p1 = g["wis_phi0"], g["wis_phi1"]
    n, Dm = p1.shape[0], p0.shape[-1]
    a0, a1 = DArray.from_numpy(p0), DArray.from_numpy(p1)
    u0, u1, cg = DArray.empty(p0.shape), DArray.empty((n,)), DArray.empty((n,))
    be.wishart_moments(a0.ptr, a1.ptr, n, n, Dm, u0.ptr, u1.ptr, cg.ptr, True)
    close(u0.numpy(), g["wis_u0"]); close(u1.numpy(), g["wis_u1"]); close(cg.numpy(), g["wis_g"])
    # Dirichlet
    p = g["dir_phi0"]
    a = DArray.from_numpy(p)
    u, cg = DArray.empty(p.shape), DArray.empty((p.shape[0],))
    be.dirichlet_moments(a.ptr, p.shape[0], p.shape[1], u.ptr, cg.ptr, True)
    close(u.numpy(), g["dir_u0"]); close(cg.numpy(), g["dir_g"])
    # Categorical softmax
    p = g["cat_phi0"]
    a = DArray.from_numpy(p)
    u, cg = DArray.empty(p.shape), DArray.empty((p.shape[0],))
    be.softmax_moments(a.ptr, p.shape[0], p.shape[1], u.ptr, cg.ptr)
    close(u.numpy(), g["cat_u0"]); close(cg.numpy(), g["cat_g"])
    # one-hot is index work: bit-exact
    lab = np.array([0, 3, 5, 5, 1], dtype=np.int64)
    l = DArray.from_numpy(lab, "i8")
    oh = DArray.empty((5, 6))
    be.one_hot(l.ptr, 5, 6, oh.ptr, True)
    assert np.array_equal(oh.numpy(), np.eye(6)[lab])
    with pytest.raises(ValueError):
        bad = DArray.from_numpy(np.array([0, 6], dtype=np.int64), "i8")
        be.one_hot(bad.ptr, 2, 6, DArray.empty((2, 6)).ptr, True)


@pytest.mark.parametrize("M,N,K", [(64, 64, 64), (70, 130, 33), (256, 1000, 1024), (9, 3000, 17), (300, 8, 128),
                                   (64, 200, 20000), (130, 70, 9001)])     # the last two: few output tiles, long k -> split-K
def test_sum_product_gemm_shapes(backend, M, N, K):
    """Contractions that collapse to C[m,n] = sum_k A[m,k] B[k,n] (dot.py:403,581 patterns) take the tensor-pipe
    GEMM path of bpk_sum_multiply on the GPU: every operand layout (k-contiguous / m-contiguous, transposed output),
    scale, accumulate, ragged tiles."""
    from bayespy_b200 import darray as D
    if M * N * K > 5e6 and backend.name != "cuda":
        pytest.skip("large contractions run on the GPU only (the oracle's einsum takes minutes)")
    rs = np.random.RandomState(M + N + K)
    a, b = rs.randn(M, K), rs.randn(N, K)
    A, B = DArray.from_numpy(a), DArray.from_numpy(b)
    ref = a @ b.T
    out = D.sum_product([A, B], [["m", "k"], ["n", "k"]], ["m", "n"])                       # both k-contiguous
    np.testing.assert_allclose(out.numpy(), ref, rtol=1e-11, atol=1e-10)
    At, Bt = DArray.from_numpy(np.ascontiguousarray(a.T)), DArray.from_numpy(np.ascontiguousarray(b.T))
    out = D.sum_product([At, Bt], [["k", "m"], ["k", "n"]], ["n", "m"], scale=-0.5)         # m-/n-contiguous, C^T
    np.testing.assert_allclose(out.numpy(), -0.5 * ref.T, rtol=1e-11, atol=1e-10)
    acc = DArray.from_numpy(np.ones((M, N)))
    D.sum_product([A, Bt], [["m", "k"], ["k", "n"]], ["m", "n"], out=acc, accumulate=True)  # accumulate onto ones
    np.testing.assert_allclose(acc.numpy(), 1.0 + ref, rtol=1e-11, atol=1e-10)
    # the <f f> pattern of SumMultiply: sum_ij cc[m,1,i,j] xx[n,i,j]
    if K == 64:
        cc, xx = rs.randn(M, 1, 8, 8), rs.randn(N, 8, 8)
        out = D.sum_product([DArray.from_numpy(cc), DArray.from_numpy(xx)],
                            [["m", "o", "i", "j"], ["n", "i", "j"]], ["m", "n"])
        np.testing.assert_allclose(out.numpy(), np.einsum("moij,nij->mn", cc, xx), rtol=1e-11, atol=1e-10)
