"""TEST INFRASTRUCTURE / CPU BASELINE — NumPy/SciPy port of the reference's VB sweep for
the Bayesian PCA model of doc/source/examples/pca.rst:40-66.

It restates, node by node and with the same kinds of NumPy/SciPy calls and the same
materialised temporaries ((1,N,K,K) second moments, (M,N) <f>,<f^2>, np.where masking,
per-plate SciPy Cholesky loops when the mask makes the precisions per-plate), what
``Q.update()`` does in the reference:

    X.update / C.update   dot.py:425-633 (einsum messages) + gaussian.py:649-706
    alpha.update          gaussian.py:609-637 (index 0 chain) + gamma.py:116-148
    tau.update            dot.py:316-415 (<f>,<f^2>) + gaussian.py:2361-2369 + gamma.py
    lower bound           expfamily.py:400-480 for Y, X, C, alpha, tau (vmp.py:192-199)

It is pinned against the reference's own output by tests/test_oracle_models.py
(golden vectors from tests/golden/make_golden.py) and is used (a) as the checker of
the CUDA path and (b) as bench.py's ``cpu_baseline`` / ``--impl reference`` arm
(kind "port"), because the pure-Python reference cannot travel to the GPU box.
Only tests/, __graft_entry__.smoke() and bench.py may import this module.
"""
import numpy as np
import scipy.linalg
import scipy.special as sp

LOG2PI = np.log(2 * np.pi)


# ---- linalg.py:31-223 restated: Python loop over plates, one LAPACK call per matrix ----------
def chol(C):
    U = np.empty(np.shape(C))
    for i in np.ndindex(*np.shape(C)[:-2]):
        U[i] = scipy.linalg.cho_factor(C[i])[0]
    return U


def chol_solve(U, b):
    """b: (..., K) vectors; plates of U and b broadcast (loop over U's plates, linalg.py:111-146)."""
    pu, pb = U.shape[:-2], b.shape[:-1]
    sh = np.broadcast_shapes(pu, pb)
    out = np.zeros(sh + b.shape[-1:])
    bb = np.broadcast_to(b, sh + b.shape[-1:]) if len(pb) < len(sh) else b
    for i in np.ndindex(*pu):
        sel = tuple(i[j] if (bb.shape[j] == pu[j]) else slice(None) for j in range(len(pu)))
        rhs = bb[sel]
        shape = rhs.shape
        rhs2 = rhs.reshape((-1, shape[-1]))
        out[sel] = scipy.linalg.cho_solve((U[i], False), rhs2.T).T.reshape(shape)
    return out


def chol_inv(U):
    V = np.tile(np.identity(U.shape[-1]), U.shape[:-2] + (1, 1))
    for i in np.ndindex(*U.shape[:-2]):
        V[i] = scipy.linalg.cho_solve((U[i], False), V[i])
    return V


def chol_logdet(U):
    return 2 * np.sum(np.log(np.einsum('...ii->...i', U)), axis=-1)


def gaussian_moments(phi0, phi1):
    """GaussianARDDistribution.compute_moments_and_cgf, gaussian.py:672-706."""
    L = chol(-2 * phi1)
    Cov = chol_inv(L)
    u0 = chol_solve(L, phi0)
    u1 = u0[..., :, None] * u0[..., None, :] + Cov
    g = -0.5 * np.einsum('...i,...i', u0, phi0) + 0.5 * chol_logdet(L)
    return u0, u1, g


def gamma_moments(phi0, phi1):
    """gamma.py:124-148."""
    log_b = np.log(-phi0)
    u0 = phi1 / (-phi0)
    u1 = sp.digamma(phi1) - log_b
    g = phi1 * log_b - sp.gammaln(phi1)
    return u0, u1, g


class PcaOracle:
    """State and sweep of  Y=GaussianARD(SumMultiply('d,d->',X,C), tau)  with
    X=GaussianARD(0,1,plates=(1,N),shape=(K,)), C=GaussianARD(0,alpha,plates=(M,1),shape=(K,)),
    alpha=Gamma(a0,b0,plates=(K,)), tau=Gamma(a0,b0)."""

    def __init__(self, y, K, C_init, mask=None, a0=1e-5, b0=1e-5):
        self.y = np.asarray(y, dtype=np.float64)
        self.M, self.N = self.y.shape
        self.K = K
        self.a0, self.b0 = a0, b0
        self.mask = True if mask is None else np.asarray(mask, dtype=bool)
        M, N = self.M, self.N
        # initialize_from_prior of every node (expfamily.py:168-180)
        self.alpha_phi = [np.full((K,), -b0), np.full((1,), a0)]
        self.alpha_u0, self.alpha_u1, self.alpha_g = gamma_moments(-b0 * np.ones(K), a0 * np.ones(K))
        self.tau_phi = [np.array(-b0), np.array(a0)]
        self.tau_u0, self.tau_u1, self.tau_g = gamma_moments(np.array(-b0), np.array(a0))
        self.X_phi = [np.zeros((1, 1, K)), (-0.5 * np.identity(K))[None, None]]
        x0, x1, xg = gaussian_moments(*self.X_phi)
        self.X_u0 = np.broadcast_to(x0, (1, 1, K)).copy()
        self.X_u1 = x1.copy()
        self.X_g = xg
        # C.initialize_from_value(C_init): u = [x, xx^T], g = inf (expfamily.py:183-206)
        c = np.asarray(C_init, dtype=np.float64).reshape(M, 1, K)
        self.C_u0 = c.copy()
        self.C_u1 = c[..., :, None] * c[..., None, :]
        self.C_g = np.inf
        self.C_phi = [np.zeros((1, 1, K)), np.zeros((1, 1, K, K))]
        self.y2 = self.y ** 2
        self.L = []
        self.l = []

    # ---- messages -----------------------------------------------------------------------------
    def _msg_Y_to_F(self):
        """[tau*y, -tau/2] masked (gaussian.py:609-637, :2351-2360, node.py:650)."""
        m0 = np.where(self.mask, self.tau_u0 * self.y, 0)
        m1 = np.where(self.mask, -0.5 * self.tau_u0, 0)
        return m0, m1

    def _F_moments(self):
        """<f>, <f^2> over the full (M,N) grid (dot.py:355,403)."""
        f = np.einsum('onk,mok->mn', self.X_u0, self.C_u0)
        ff = np.einsum('onkl,mokl->mn', np.broadcast_to(self.X_u1, (1, self.N, self.K, self.K))
                       if self.X_u1.shape[1] == 1 else self.X_u1, self.C_u1)
        return f, ff

    def update_X(self):
        K, N = self.K, self.N
        m0, m1 = self._msg_Y_to_F()
        msg0 = np.einsum('mok,mn->onk', self.C_u0, m0)                       # dot.py:581 (ind 0)
        if np.ndim(m1) == 0:
            msg1 = (m1 * np.einsum('mokl->kl', self.C_u1))[None, None]         # collapses to (1,1,K,K)
        else:
            msg1 = np.einsum('mokl,mn->onkl', self.C_u1, m1)                   # dot.py:581 (ind 1)
        phi0 = np.zeros((1, 1, K)) + msg0                                      # prior mu=0, alpha=1
        phi1 = (-0.5 * np.identity(K))[None, None] + msg1
        self.X_phi = [phi0, phi1]
        u0, u1, g = gaussian_moments(phi0, phi1)
        self.X_u0 = np.array(np.broadcast_to(u0, (1, N, K)))                  # _set_moments copies
        self.X_u1 = np.array(np.broadcast_to(u1, (1, N, K, K)))
        self.X_g = g

    def update_C(self):
        K, M = self.K, self.M
        m0, m1 = self._msg_Y_to_F()
        msg0 = np.einsum('onk,mn->mok', self.X_u0, m0)
        if np.ndim(m1) == 0:
            msg1 = (m1 * np.einsum('onkl->kl', self.X_u1))[None, None]
        else:
            msg1 = np.einsum('onkl,mn->mokl', self.X_u1, m1)
        a = self.alpha_u0
        phi0 = np.zeros((1, 1, K)) + msg0
        phi1 = (-0.5 * np.diag(a))[None, None] + msg1
        self.C_phi = [phi0, phi1]
        u0, u1, g = gaussian_moments(phi0, phi1)
        self.C_u0 = np.array(np.broadcast_to(u0, (M, 1, K)))
        self.C_u1 = np.array(np.broadcast_to(u1, (M, 1, K, K)))
        self.C_g = g

    def update_alpha(self):
        """message [-1/2 <c_k^2>, 1/2] summed over M (gaussian.py:627-635, :2361-2369)."""
        c2 = np.einsum('mokk->mok', self.C_u1)
        m0 = np.sum(-0.5 * c2, axis=(0, 1))
        m1 = 0.5 * self.M * np.ones(self.K)
        self.alpha_phi = [-self.b0 + m0, self.a0 + m1]
        self.alpha_u0, self.alpha_u1, self.alpha_g = gamma_moments(*self.alpha_phi)

    def _e2(self):
        f, ff = self._F_moments()
        return np.where(self.mask, self.y2 - 2 * self.y * f + ff, 0)

    def update_tau(self):
        e2 = self._e2()
        nobs = self.M * self.N if self.mask is True else np.sum(self.mask)
        m0 = -0.5 * np.sum(e2)
        m1 = 0.5 * nobs
        self.tau_phi = [np.array(-self.b0 + m0), np.array(self.a0 + m1)]
        self.tau_u0, self.tau_u1, self.tau_g = gamma_moments(*self.tau_phi)

    # ---- lower bound (expfamily.py:400-480 per node) ---------------------------------------------------
    def lower_bound(self):
        K, M, N = self.K, self.M, self.N
        tau, logtau = self.tau_u0, self.tau_u1
        # Y (observed): cgf + f + phi.u, masked
        f, ff = self._F_moments()
        LY = (-0.5 * tau * ff + 0.5 * logtau) - 0.5 * LOG2PI + tau * f * self.y - 0.5 * tau * self.y2
        lY = np.sum(np.where(self.mask, LY, 0))
        # X (latent): prior N(0, I)
        phi_p0, phi_p1 = np.zeros((1, 1, K)), (-0.5 * np.identity(K))[None, None]
        LX = 0.0 - self.X_g + np.sum((phi_p0 - self.X_phi[0]) * self.X_u0, axis=-1) \
            + np.sum((phi_p1 - self.X_phi[1]) * self.X_u1, axis=(-1, -2))
        lX = np.sum(LX) * (N / LX.shape[1] if LX.shape[1] != N else 1)
        # C (latent): prior N(0, diag(alpha)^-1)
        a, loga = self.alpha_u0, self.alpha_u1
        phi_p1 = (-0.5 * np.diag(a))[None, None]
        cgf = 0.5 * np.sum(loga)
        C_phi0 = self.C_phi[0]
        LC = cgf - self.C_g + np.sum((0 - C_phi0) * self.C_u0, axis=-1) \
            + np.sum((phi_p1 - self.C_phi[1]) * self.C_u1, axis=(-1, -2))
        lC = np.sum(np.broadcast_to(LC, (M, 1)))
        # alpha, tau (latent gamma): cgf_p = a0 log b0 - lgamma(a0)
        cg = self.a0 * np.log(self.b0) - sp.gammaln(self.a0)
        La = cg - self.alpha_g + (-self.b0 - self.alpha_phi[0]) * self.alpha_u0 \
            + (self.a0 - self.alpha_phi[1]) * self.alpha_u1
        la = np.sum(La)
        lt = cg - self.tau_g + (-self.b0 - self.tau_phi[0]) * self.tau_u0 \
            + (self.a0 - self.tau_phi[1]) * self.tau_u1
        terms = dict(Y=float(lY), X=float(lX), C=float(lC), alpha=float(la), tau=float(lt))
        return float(lY + lX + lC + la + lt), terms

    def sweep(self, order=("X", "C", "alpha", "tau")):
        """One Q.update() iteration: node updates in model order, then the bound (vmp.py:154-172)."""
        for name in order:
            getattr(self, "update_" + name)()
        L, terms = self.lower_bound()
        self.L.append(L)
        self.l.append(terms)
        return L


def make_data(M, N, seed=1, K_true=4):
    """Synthetic inputs of SURVEY §8d / demos/pca.py:68-71 (NumPy legacy RNG)."""
    np.random.seed(seed)
    w = np.random.randn(M, K_true)
    x = np.random.randn(N, K_true)
    return w @ x.T + 0.1 * np.random.randn(M, N)
