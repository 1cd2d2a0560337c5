"""Fused sweep kernels (bpk_pca_xsweep / bpk_pca_stats / bpk_pca_xsweep_masked / bpk_gmm_sweep /
bpk_sumsq) against plain NumPy on seeded inputs, through the C-ABI wrappers.
Sizes cover: ragged tiles, odd N (8-byte cp.async path), M<64 / K<16 padding, the generic
M>64 / K>16 path, and one L2-exceeding size with size-independent properties."""
import numpy as np
import pytest

from bayespy_b200.darray import DArray


def _pca_ref(y, A, b):
    x = y.T @ A.T + (b if b is not None else 0.0)
    return x, y @ x, x.T @ x, x.sum(0)


@pytest.mark.parametrize("M,N,K", [(64, 1000, 16), (64, 129, 16), (64, 4097, 16), (20, 100, 5), (8, 7, 3),
                                   (64, 2, 16), (33, 1023, 9), (80, 300, 16), (64, 300, 20), (1, 50, 1)])
@pytest.mark.parametrize("with_b", [False, True])
def test_pca_xsweep(backend, M, N, K, with_b):
    rng = np.random.RandomState(M * 1000 + N + K)
    y = rng.randn(M, N)
    A = rng.randn(K, M) / np.sqrt(M)
    b = rng.randn(K) if with_b else None
    x0, syx, sxx, sx = _pca_ref(y, A, b)
    Y, Ad = DArray.from_numpy(y), DArray.from_numpy(A)
    bd = DArray.from_numpy(b) if with_b else None
    X = DArray.empty((N, K))
    st = DArray.zeros((M * K + K * K + K,))
    backend.pca_xsweep(Y.ptr, M, N, K, Ad.ptr, bd.ptr if with_b else 0, X.ptr, st.ptr)
    s = st.numpy()
    np.testing.assert_allclose(X.numpy(), x0, rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(s[:M * K].reshape(M, K), syx, rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(s[M * K:M * K + K * K].reshape(K, K), sxx, rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(s[M * K + K * K:], sx, rtol=1e-10, atol=1e-9)
    # statistics-only entry point accumulates onto what is there
    backend.pca_stats(Y.ptr, M, N, K, X.ptr, st.ptr)
    np.testing.assert_allclose(st.numpy(), 2 * s, rtol=1e-10, atol=1e-9)


@pytest.mark.gpu
def test_pca_xsweep_full_size_properties(cuda_backend):
    """N=2e6 (1 GB of Y, > L2): size-independent checks — linearity in A, consistency of the
    statistics with the written X (checksum of checksums), and agreement of the two entry points."""
    be = cuda_backend
    M, N, K = 64, 2_000_000, 16
    rng = np.random.default_rng(5)
    y = rng.standard_normal((M, N))
    A = rng.standard_normal((K, M)) / 8
    Y, Ad, A2 = DArray.from_numpy(y), DArray.from_numpy(A), DArray.from_numpy(2 * A)
    X, X2 = DArray.empty((N, K)), DArray.empty((N, K))
    st, st2, st3 = (DArray.zeros((M * K + K * K + K,)) for _ in range(3))
    be.pca_xsweep(Y.ptr, M, N, K, Ad.ptr, 0, X.ptr, st.ptr)
    be.pca_xsweep(Y.ptr, M, N, K, A2.ptr, 0, X2.ptr, st2.ptr)
    be.pca_stats(Y.ptr, M, N, K, X.ptr, st3.ptr)
    x, x2, s, s2, s3 = X.numpy(), X2.numpy(), st.numpy(), st2.numpy(), st3.numpy()
    np.testing.assert_allclose(x2, 2 * x, rtol=1e-13)
    np.testing.assert_allclose(s2[:M * K], 2 * s[:M * K], rtol=1e-12)
    np.testing.assert_allclose(s2[M * K:M * K + K * K], 4 * s[M * K:M * K + K * K], rtol=1e-12)
    np.testing.assert_allclose(s3, s, rtol=1e-11, atol=1e-6)
    np.testing.assert_allclose(s[M * K + K * K:], x.sum(0), rtol=1e-9, atol=1e-6)
    np.testing.assert_allclose(s[M * K:M * K + K * K].reshape(K, K), x.T @ x, rtol=1e-10)
    sub = slice(0, 50_000)
    np.testing.assert_allclose(x[sub], y[:, sub].T @ A.T, rtol=1e-11, atol=1e-11)


@pytest.mark.parametrize("count,masked", [(1000, False), (12345, True), (1, False)])
def test_sumsq(backend, count, masked):
    rng = np.random.RandomState(count)
    y = rng.randn(count)
    m = rng.rand(count) > 0.3
    Y = DArray.from_numpy(y)
    md = DArray.from_numpy(m)
    out = DArray.empty((2,))
    backend.sumsq(Y.ptr, md.ptr if masked else 0, count, out.ptr)
    o = out.numpy()
    if masked:
        np.testing.assert_allclose(o, [np.sum(y[m] ** 2), m.sum()], rtol=1e-12)
    else:
        np.testing.assert_allclose(o, [np.sum(y ** 2), count], rtol=1e-12)


@pytest.mark.parametrize("M,N,K", [(12, 40, 4), (64, 130, 16), (5, 33, 7)])
def test_pca_xsweep_masked(backend, M, N, K):
    rng = np.random.RandomState(N)
    y = rng.randn(M, N)
    mask = rng.rand(M, N) > 0.25
    W = rng.randn(M, K)
    Cw = rng.randn(M, K, K)
    WW = W[:, :, None] * W[:, None, :] + 0.1 * Cw @ np.swapaxes(Cw, -1, -2)
    tau, alpha, amu = 1.7, rng.gamma(2.0, 1.0, K) + 0.1, rng.randn(K)
    mk = mask.astype(float)
    Lam = np.diag(alpha)[None] + tau * np.einsum("mn,mij->nij", mk, WW)
    phi0 = tau * np.einsum("mn,mn,mk->nk", mk, y, W) + amu
    cov = np.linalg.inv(Lam)
    x = np.einsum("nij,nj->ni", cov, phi0)
    g = -0.5 * np.einsum("ni,ni->n", x, phi0) + 0.5 * np.linalg.slogdet(Lam)[1]
    xx = cov + x[:, :, None] * x[:, None, :]
    d = {k: DArray.from_numpy(v) for k, v in dict(y=y, W=W, WW=WW, alpha=alpha, amu=amu).items()}
    md = DArray.from_numpy(mask)
    X, COV, G = DArray.empty((N, K)), DArray.empty((N, K, K)), DArray.empty((N,))
    st = DArray.zeros((M * K + M * K * K,))
    backend.pca_xsweep_masked(d["y"].ptr, md.ptr, M, N, K, d["W"].ptr, d["WW"].ptr, tau, d["alpha"].ptr,
                              d["amu"].ptr, X.ptr, COV.ptr, G.ptr, st.ptr, True)
    np.testing.assert_allclose(X.numpy(), x, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(COV.numpy(), cov, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(G.numpy(), g, rtol=1e-10)
    s = st.numpy()
    np.testing.assert_allclose(s[:M * K].reshape(M, K), np.einsum("mn,mn,nk->mk", mk, y, x), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(s[M * K:].reshape(M, K, K), np.einsum("mn,nij->mij", mk, xx), rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("M,N,K,chunk_tiles", [(12, 40, 4, None), (64, 130, 16, None), (5, 33, 7, None), (64, 40000, 16, 1),
                                               (64, 19001, 16, 1), (33, 1025, 16, None)])
def test_pca_xsweep_masked_fused(backend, M, N, K, chunk_tiles, monkeypatch):
    """The fused masked sweep (csrc/pca_masked.cu): masked tensor-pipe GEMMs + one thread per column, several chunks
    of the fixed-size scratch, ragged last tile, padded M and K — against a dense NumPy restatement."""
    if N > 5000 and backend.name != "cuda":
        pytest.skip("multi-chunk sizes run on the GPU only (the dense NumPy restatement is the checker there)")
    if chunk_tiles is not None:
        monkeypatch.setenv("BPK_PMASK_CHUNK_TILES", str(chunk_tiles))      # 148 x 128 columns per chunk
    rng = np.random.RandomState(N % 1000)
    y = rng.randn(M, N)
    mask = rng.rand(M, N) > 0.25
    W = rng.randn(M, K)
    Cw = rng.randn(M, K, K)
    WW = W[:, :, None] * W[:, None, :] + 0.1 * Cw @ np.swapaxes(Cw, -1, -2)
    tau, alpha, amu = 1.7, rng.gamma(2.0, 1.0, K) + 0.1, rng.randn(K)
    mk = mask.astype(float)
    Lam = np.diag(alpha)[None] + tau * np.einsum("mn,mij->nij", mk, WW)
    phi0 = tau * np.einsum("mn,mn,mk->nk", mk, y, W) + amu
    cov = np.linalg.inv(Lam)
    x = np.einsum("nij,nj->ni", cov, phi0)
    q = np.einsum("ni,ni->n", x, phi0)
    ld = np.linalg.slogdet(Lam)[1]
    xx = cov + x[:, :, None] * x[:, None, :]
    d = {k: DArray.from_numpy(v) for k, v in dict(y=y, W=W, WW=WW, alpha=alpha, amu=amu).items()}
    md = DArray.from_numpy(mask)
    X, G = DArray.empty((N, K)), DArray.empty((N,))
    st = DArray.zeros((M * K + M * K * K + K * K + K + 2,))
    backend.pca_xsweep_masked_fused(d["y"].ptr, md.ptr, M, N, K, d["W"].ptr, d["WW"].ptr, tau, d["alpha"].ptr,
                                    d["amu"].ptr, X.ptr, G.ptr, st.ptr, True)
    np.testing.assert_allclose(X.numpy(), x, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(G.numpy(), -0.5 * q + 0.5 * ld, rtol=1e-10, atol=1e-10)
    s = st.numpy()
    o = 0
    np.testing.assert_allclose(s[o:o + M * K].reshape(M, K), np.einsum("mn,mn,nk->mk", mk, y, x), rtol=1e-9, atol=1e-8)
    o += M * K
    np.testing.assert_allclose(s[o:o + M * K * K].reshape(M, K, K), np.einsum("mn,nij->mij", mk, xx), rtol=1e-9, atol=1e-8)
    o += M * K * K
    np.testing.assert_allclose(s[o:o + K * K].reshape(K, K), xx.sum(axis=0), rtol=1e-9, atol=1e-8)
    o += K * K
    np.testing.assert_allclose(s[o:o + K], x.sum(axis=0), rtol=1e-9, atol=1e-8)
    np.testing.assert_allclose(s[o + K], q.sum(), rtol=1e-10)
    np.testing.assert_allclose(s[o + K + 1], ld.sum(), rtol=1e-10)
    # a second call ADDS to the statistics (the caller zeroes them)
    backend.pca_xsweep_masked_fused(d["y"].ptr, md.ptr, M, N, K, d["W"].ptr, d["WW"].ptr, tau, d["alpha"].ptr,
                                    d["amu"].ptr, X.ptr, G.ptr, st.ptr, True)
    np.testing.assert_allclose(st.numpy()[o + K], 2 * q.sum(), rtol=1e-10)


def test_pca_xsweep_masked_fused_rejects_a_non_spd_precision(backend):
    M, N, K = 8, 20, 4
    rng = np.random.RandomState(0)
    W = rng.randn(M, K)
    WW = -np.tile(np.identity(K), (M, 1, 1))                     # "second moments" that make Lam_n indefinite
    d = {k: DArray.from_numpy(v) for k, v in dict(y=rng.randn(M, N), W=W, WW=WW, alpha=np.ones(K), amu=np.zeros(K)).items()}
    md = DArray.from_numpy(np.ones((M, N), dtype=bool))
    X, G = DArray.empty((N, K)), DArray.empty((N,))
    st = DArray.zeros((M * K + M * K * K + K * K + K + 2,))
    with pytest.raises(Exception, match="positive definite"):
        backend.pca_xsweep_masked_fused(d["y"].ptr, md.ptr, M, N, K, d["W"].ptr, d["WW"].ptr, 1.0, d["alpha"].ptr,
                                        d["amu"].ptr, X.ptr, G.ptr, st.ptr, True)


@pytest.mark.parametrize("variant", [None, "BPK_GMM_V1", "BPK_GMM_V3", "BPK_GMM_V0"])
@pytest.mark.parametrize("N,D,K", [(300, 3, 5), (1000, 8, 64), (129, 2, 10), (77, 1, 2), (70001, 8, 64)])
def test_gmm_sweep(backend, N, D, K, variant, monkeypatch):
    """Every variant of the mixture sweep kernel (default: warp pairs; V1: CTA barriers; V3: warp quads of 128
    registers; V0: scalar) against the dense NumPy restatement."""
    if variant is not None:
        if backend.name != "cuda":
            pytest.skip("kernel variants exist on the GPU only")
        monkeypatch.setenv(variant, "1")
    if N > 5000 and backend.name != "cuda":
        pytest.skip("large case runs on the GPU only")
    rng = np.random.RandomState(N + D)
    y = 3 * rng.randn(N, D)
    c = rng.randn(K)
    h = rng.randn(K, D)
    R = rng.randn(K, D, D)
    Lam = R @ np.swapaxes(R, -1, -2) + np.identity(D)
    logpi = np.log(rng.dirichlet(np.ones(K)))
    L = c[None] + y @ h.T - 0.5 * np.einsum("ni,kij,nj->nk", y, Lam, y) + logpi[None]
    m = L.max(-1, keepdims=True)
    lse = np.log(np.exp(L - m).sum(-1, keepdims=True)) + m
    p = np.exp(L - lse)
    p /= p.sum(-1, keepdims=True)
    d = {k: DArray.from_numpy(v) for k, v in dict(y=y, c=c, h=h, Lam=Lam, logpi=logpi).items()}
    P, G = DArray.empty((N, K)), DArray.empty((N,))
    st = DArray.zeros((K + K * D + K * D * D + 1,))
    backend.gmm_sweep(d["y"].ptr, N, D, K, d["c"].ptr, d["h"].ptr, d["Lam"].ptr, d["logpi"].ptr, P.ptr, G.ptr, st.ptr)
    np.testing.assert_allclose(P.numpy(), p, rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(G.numpy(), -lse[:, 0], rtol=1e-11)
    s = st.numpy()
    np.testing.assert_allclose(s[:K], p.sum(0), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(s[K:K + K * D].reshape(K, D), p.T @ y, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(s[K + K * D:-1].reshape(K, D, D), np.einsum("nk,ni,nj->kij", p, y, y), rtol=1e-10,
                               atol=1e-10)
    np.testing.assert_allclose(s[-1], lse.sum(), rtol=1e-11)


@pytest.mark.parametrize("N,D,K", [(300, 3, 5), (1000, 8, 64)])
def test_gmm_stats(backend, N, D, K):
    rng = np.random.RandomState(N)
    y = rng.randn(N, D)
    p = rng.dirichlet(np.ones(K), size=N)
    Y, P = DArray.from_numpy(y), DArray.from_numpy(p)
    st = DArray.zeros((K + K * D + K * D * D + 1,))
    backend.gmm_stats(Y.ptr, N, D, K, P.ptr, st.ptr)
    s = st.numpy()
    np.testing.assert_allclose(s[:K], p.sum(0), rtol=1e-11)
    np.testing.assert_allclose(s[K:K + K * D].reshape(K, D), p.T @ y, rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(s[K + K * D:-1].reshape(K, D, D), np.einsum("nk,ni,nj->kij", p, y, y), rtol=1e-10,
                               atol=1e-11)
    assert s[-1] == 0.0
