"""Rotation parameter expansion: ``RotationOptimizer`` and ``RotateGaussianARD`` of
bayespy/inference/vmp/transformations.py:23-222 and :376-1110, for the case every documented PCA / factor-model
example uses: rotation of the variable axis of a ``GaussianARD`` block (``axis=-1``, no plate rotation, no subset),
with its ARD precision either updated (``RotateGaussianARD(C, alpha)``) or fixed (``RotateGaussianARD(X)``).

Split of work: the plate-summed second moments the cost function needs (sum <x x^T>, sum <x> mu^T — K x K each) are
reduced on the device by the same plate-sum kernels the VB sweep uses and read back once per ``rotate()``; the
optimisation over the K x K matrix R (SciPy CG, exactly the reference's call) runs on the host; applying R to the
plated moments (u0 <- u0 R^T over all N columns, Cov <- R Cov R^T) is again device work.
"""
import warnings

import numpy as np
from scipy import optimize

from ... import darray as D
from ...darray import DArray
from ...engine.gaussian import FactoredSecondMoment, GaussianARD, dense
from ...engine.gamma import Gamma
from ...engine.plans import LazyArray


def _np(a):
    return np.asarray(a.numpy() if hasattr(a, "numpy") else a, dtype=np.float64)


class RotateGaussianARD:
    """``RotateGaussianARD(X)`` / ``RotateGaussianARD(C, alpha)`` (transformations.py:376-1110)."""

    def __init__(self, X, *alpha, axis=-1, precompute=False, subset=None):
        if not isinstance(X, GaussianARD) or len(X.dims[0]) != 1:
            raise NotImplementedError("Rotation is implemented for GaussianARD nodes with one variable axis")
        if axis not in (-1, 0) or subset is not None or precompute:
            raise NotImplementedError("Only axis=-1 without subset / precompute is implemented")
        if len(alpha) > 1:
            raise ValueError("Too many arguments")
        self.node_X = X
        self.node_alpha = alpha[0] if alpha else None
        self.update_alpha = bool(alpha)
        if self.update_alpha and not (isinstance(self.node_alpha, Gamma) and tuple(self.node_alpha.plates) == tuple(X.dims[0])):
            raise NotImplementedError("The ARD node must be a Gamma node with plates equal to the rotated axis")
        self.D = X.dims[0][0]
        self.plate_axis = None

    def nodes(self):
        return [self.node_X, self.node_alpha] if self.update_alpha else [self.node_X]

    # ---- statistics (transformations.py:482-640, the branch without plate rotation) -------------------------
    def setup(self, plate_axis=None):
        if plate_axis is not None:
            raise NotImplementedError("Plate rotation is not implemented")
        X, K = self.node_X, self.D
        Np = int(np.prod(X.plates, dtype=np.int64)) if X.plates else 1
        u_mu, u_al = X.parents[0].get_moments(), X.parents[1].get_moments()
        x = X.u[0].reshape((-1, K))
        # sum over plates of <x x^T>
        u1 = X.u[1]
        if isinstance(u1, FactoredSecondMoment):
            XX = D.sum_product([x, x], [["n", "i"], ["n", "j"]], ["i", "j"])
            cov = u1.cov.reshape((-1, K, K))
            D.sum_product([cov], [["c", "i", "j"]], ["i", "j"], out=XX, accumulate=True, scale=float(Np) / cov.shape[0])
        else:
            XX = D.sum_product([dense(u1).reshape((-1, K, K))], [["n", "i", "j"]], ["i", "j"])
        self.XX = _np(XX)
        # prior mean, broadcast over the plates of X: sum_n <x_n> mu_n^T and sum_n <mu_n^2>
        mu = D.asarray(u_mu[0])
        mu2 = D.asarray(u_mu[1])
        full = tuple(X.plates) + (K,)
        if mu.size == 1 or tuple(mu.shape) == (K,) or all(n == 1 for n in mu.shape[:-1]):
            m = np.broadcast_to(_np(mu).reshape(-1)[-K:] if mu.size >= K else _np(mu).reshape(()), (K,))
            m2 = np.broadcast_to(_np(mu2).reshape(-1)[-K:] if mu2.size >= K else _np(mu2).reshape(()), (K,))
            sx = _np(D.sum_product([x], [["n", "i"]], ["i"]))
            self.Xmu = np.outer(sx, m)
            self.mu2 = Np * m2
        else:
            mb = mu.broadcast_to(full).reshape((-1, K))
            self.Xmu = _np(D.sum_product([x, mb], [["n", "i"], ["n", "j"]], ["i", "j"]))
            self.mu2 = _np(D.sum_product([mu2.broadcast_to(full).reshape((-1, K))], [["n", "i"]], ["i"]))
        self.Np = Np
        if self.update_alpha:
            al = self.node_alpha
            self.a = np.broadcast_to(_np(al.phi[1]), (K,)).copy()
            self.a0 = np.broadcast_to(_np(al.parents[0].get_moments()[0]), (K,)).copy()
            self.b0 = np.broadcast_to(_np(al.parents[1].get_moments()[0]), (K,)).copy()
        else:
            a = _np(u_al[0])
            if a.size not in (1, K) or (a.ndim > 1 and any(n != 1 for n in a.shape[:-1])):
                raise NotImplementedError("A fixed precision that varies over plates is not supported by the rotation")
            self.alpha = np.broadcast_to(a.reshape(-1)[-K:] if a.size >= K else a.reshape(()), (K,)).copy()

    # ---- cost function and gradient (transformations.py:642-1010) ---------------------------------------------
    def _compute_bound(self, R, logdet=None, inv=None, gradient=False, terms=False):
        """Change of the lower bound when q(x) -> q(R x) (and q(alpha) is re-optimised), and d/dR of it.

        With S = sum <x x^T>, T = sum <x> mu^T and m2 = sum mu^2 (plate sums from ``setup``), the expected squared
        deviation of the rotated variable along output dimension i is
            e_i(R) = (R S R^T)_ii - 2 (R T)_ii + m2_i,        d e_i / d R_i: = 2 (R S)_i: - 2 T_:i .
        Fixed precision a_i:   bound = -1/2 sum_i a_i e_i + n log|det R|.
        ARD precision Gamma(a0, b0) with posterior shape `a`:  b_i = b0_i + e_i / 2, <alpha_i> = a_i / b_i and
            bound = -1/2 sum_i <alpha_i> e_i - n/2 sum_i log b_i - sum_i (a0_i log b_i + b0_i <alpha_i>) + n log|det R|.
        """
        S, Tm, m2, n = self.XX, self.Xmu, self.mu2, self.Np
        RS = R @ S
        e = np.einsum("ik,ik->i", RS, R) - 2.0 * np.einsum("ik,ki->i", R, Tm) + m2
        if logdet is None:
            logdet = np.linalg.slogdet(R)[1]
            inv = np.linalg.inv(R)
        if self.update_alpha:
            rate = self.b0 + 0.5 * e
            prec = self.a / rate
            logprec = -np.log(rate)
        else:
            prec = self.alpha
            logprec = np.zeros(self.D)
        entropy_gain = n * logdet
        fit = -0.5 * np.sum(prec * e) + 0.5 * n * np.sum(logprec)
        hyper = (np.sum(self.a0 * logprec) - np.sum(self.b0 * prec)) if self.update_alpha else 0.0
        if terms:
            out = {self.node_X: fit + entropy_gain}
            if self.update_alpha:
                out[self.node_alpha] = hyper
            return out
        value = fit + hyper + entropy_gain
        if not gradient:
            return value
        de = 2.0 * RS - 2.0 * Tm.T                       # row i = d e_i / d R_i:
        g_fit = -0.5 * prec[:, None] * de
        g_hyper = 0.0
        if self.update_alpha:
            drate = 0.5 * de
            dprec = (-prec / rate)[:, None] * drate        # d <alpha_i> / d R_i:
            dlogprec = -(1.0 / rate)[:, None] * drate
            g_fit = g_fit - 0.5 * e[:, None] * dprec + 0.5 * n * dlogprec
            g_hyper = self.a0[:, None] * dlogprec - self.b0[:, None] * dprec
        return value, g_fit + g_hyper + n * inv.T

    def bound(self, R, logdet=None, inv=None, Q=None):
        return self._compute_bound(R, logdet=logdet, inv=inv, gradient=True)

    def get_bound_terms(self, R, logdet=None, inv=None, Q=None):
        return self._compute_bound(R, logdet=logdet, inv=inv, gradient=False, terms=True)

    # ---- apply (transformations.py:454-468 -> gaussian.py:1693-1745) -----------------------------------------------
    def rotate(self, R, inv=None, logdet=None, Q=None):
        self.node_X.rotate(R, inv=inv, logdet=logdet)
        if self.update_alpha:
            self.node_alpha.update()


class RotationOptimizer:
    """``RotationOptimizer(block1, block2, D)``: block1 is rotated by R, block2 by R^-T (transformations.py:23-222)."""

    def __init__(self, block1, block2, D_):
        self.block1, self.block2, self.D = block1, block2, int(D_)

    def rotate(self, maxiter=10, check_gradient=False, verbose=False, check_bound=False):
        Dm = self.D

        def cost(r):
            R = np.reshape(r, (Dm, Dm))
            invR = np.linalg.inv(R)
            logdetR = np.linalg.slogdet(R)[1]
            b1, db1 = self.block1.bound(R, logdet=logdetR, inv=invR)
            b2, db2 = self.block2.bound(invR.T, logdet=-logdetR, inv=R.T)
            db2 = -invR.T @ db2.T @ invR.T
            return -(b1 + b2), np.ravel(-(db1 + db2))

        self.block1.setup()
        self.block2.setup()
        r0 = np.ravel(np.identity(Dm))
        if check_gradient:
            Rr = np.random.randn(Dm, Dm)
            num = optimize.approx_fprime(np.ravel(Rr), lambda v: cost(v)[0], 1e-6)
            ana = cost(np.ravel(Rr))[1]
            err = np.linalg.norm(ana - num) / max(np.linalg.norm(num), 1e-300)
            if err > 1e-5:
                warnings.warn("Rotation gradient has relative error %g" % err)
        cost_begin = cost(r0)[0]
        if check_bound:
            nodes = set(self.block1.nodes()) | set(self.block2.nodes())
            true_begin = {n: float(n.lower_bound_contribution()) for n in nodes}
            t1 = dict(self.block1.get_bound_terms(np.identity(Dm)))
            t1.update(self.block2.get_bound_terms(np.identity(Dm)))
        # the reference's bayespy.utils.optimize.minimize: SciPy CG with an analytic gradient
        opt = optimize.minimize(cost, r0, jac=True, method="CG", options={"disp": verbose, "maxiter": maxiter})
        r = opt.x
        cost_end = cost(r)[0]
        R = np.reshape(r, (Dm, Dm))
        invR = np.linalg.inv(R)
        logdetR = np.linalg.slogdet(R)[1]
        if check_bound:
            t2 = dict(self.block1.get_bound_terms(R, logdet=logdetR, inv=invR))
            t2.update(self.block2.get_bound_terms(invR.T, logdet=-logdetR, inv=R.T))
        self.block1.rotate(R, inv=invR, logdet=logdetR)
        self.block2.rotate(invR.T, inv=R.T, logdet=-logdetR)
        if cost_end - cost_begin > 0:
            warnings.warn("Rotation optimization made the cost function worse by %g. Probably a bug in the gradient "
                          "of the rotation functions." % (cost_end - cost_begin,))
        if check_bound:
            total = 0.0
            for n in nodes:
                true_change = float(n.lower_bound_contribution()) - true_begin[n]
                change = t2[n] - t1[n]
                total += change
                if not np.allclose(change, true_change):
                    warnings.warn("Rotation cost function is not consistent with the true lower bound for node %s. "
                                  "Bound changed %g but optimized function changed %g." % (n.name, true_change, change))
            if total < 0:
                warnings.warn("Rotation made the true lower bound worse by %g. Probably a bug in the rotation "
                              "functions." % total)
