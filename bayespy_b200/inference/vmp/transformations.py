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
        self.plate_axis = plate_axis
        if plate_axis is not None:
            return self._setup_with_plate_rotation(plate_axis)
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

    def _setup_with_plate_rotation(self, plate_axis):
        """Statistics for a joint rotation of the variable axis (by R) and of ONE plate axis (by Q), the case the
        dynamics matrix of a state-space model needs: A -> Q A R^T  (transformations.py:486-492, :574-626).  Here the
        node has exactly one plate axis (the rotated one); everything is K x K x P small and lives on the host."""
        X, K = self.node_X, self.D
        if not isinstance(plate_axis, int):
            raise ValueError("Plate axis must be integer")
        if len(X.plates) != 1 or plate_axis not in (-1, 0):
            raise NotImplementedError("Plate rotation is implemented for nodes with exactly one plate axis")
        P = X.plates[0]
        x = np.broadcast_to(_np(X.u[0]), (P, K)).copy()
        xx = np.broadcast_to(_np(dense(X.u[1])), (P, K, K))
        self.Xp = x                                                     # <x_p>
        self.CovXp = xx - x[:, :, None] * x[:, None, :]                  # Cov(x_p)
        u_mu, u_al = X.parents[0].get_moments(), X.parents[1].get_moments()
        self.mup = np.broadcast_to(_np(u_mu[0]), (P, K)).copy()
        self.mu2 = np.sum(np.broadcast_to(_np(u_mu[1]), (P, K)), axis=0)
        self.Np = P
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

    def _plate_rotated_statistics(self, Q):
        """S = sum_p <x'_p x'_p^T>, T = sum_p <x'_p> mu_p^T and the entropy term of the plate mixing for
        <x'_i> = sum_k Q_ik <x_k>, Cov(x'_p) ~ (sum_i Q_ip)^2 Cov(x_p)  — the approximation gaussian.py:1743-1774 applies
        (transformations.py:713-752)."""
        s = np.sum(Q, axis=0)
        QX = Q @ self.Xp
        S = np.einsum("p,pij->ij", s * s, self.CovXp) + QX.T @ QX
        Tm = QX.T @ self.mup
        return S, Tm, self.D * np.sum(np.log(np.abs(s))), s, QX

    # ---- cost function and gradient (transformations.py:642-1010) ---------------------------------------------
    def _compute_bound(self, R, logdet=None, inv=None, gradient=False, terms=False, Q=None):
        """Change of the lower bound when q(x) -> q(R x) (and q(alpha) is re-optimised), and d/dR of it.

        With S = sum <x x^T>, T = sum <x> mu^T and m2 = sum mu^2 (plate sums from ``setup``), the expected squared
        deviation of the rotated variable along output dimension i is
            e_i(R) = (R S R^T)_ii - 2 (R T)_ii + m2_i,        d e_i / d R_i: = 2 (R S)_i: - 2 T_:i .
        Fixed precision a_i:   bound = -1/2 sum_i a_i e_i + n log|det R|.
        ARD precision Gamma(a0, b0) with posterior shape `a`:  b_i = b0_i + e_i / 2, <alpha_i> = a_i / b_i and
            bound = -1/2 sum_i <alpha_i> e_i - n/2 sum_i log b_i - sum_i (a0_i log b_i + b0_i <alpha_i>) + n log|det R|.
        """
        plate_entropy = 0.0
        if self.plate_axis is not None:
            if Q is None:
                raise ValueError("Plates should be rotated but no Q given")
            S, Tm, plate_entropy, colsum, QX = self._plate_rotated_statistics(Q)
            m2, n = self.mu2, self.Np
        else:
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
        entropy_gain = n * logdet + plate_entropy
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
        dR = g_fit + g_hyper + n * inv.T
        if self.plate_axis is None:
            return value, dR
        # d/dQ: every term depends on Q through e_i only, plus the entropy of the plate mixing.
        #   w_i = d value / d e_i;   with Y = X R^T (rotated means) and c_pi = (R Cov_p R^T)_ii
        #   d e_i / d Q_ab = 2 s_b c_bi + 2 (Q Y)_ai Y_bi - 2 mu_ai Y_bi
        if self.update_alpha:
            w = -0.5 * prec + (0.5 * e * prec / rate - 0.5 * n / rate - self.a0 / rate + self.b0 * prec / rate) * 0.5
        else:
            w = -0.5 * prec * np.ones(self.D)
        Y = self.Xp @ R.T
        c = np.einsum("ik,pkl,il->pi", R, self.CovXp, R)
        dQ = (2.0 * colsum * (c @ w))[None, :] + 2.0 * ((QX @ R.T) * w) @ Y.T - 2.0 * (self.mup * w) @ Y.T \
            + (self.D / colsum)[None, :]
        return value, dR, dQ * np.ones((self.Np, 1))

    def bound(self, R, logdet=None, inv=None, Q=None):
        return self._compute_bound(R, logdet=logdet, inv=inv, gradient=True, Q=Q)

    def get_bound_terms(self, R, logdet=None, inv=None, Q=None):
        return self._compute_bound(R, logdet=logdet, inv=inv, gradient=False, terms=True, Q=Q)

    # ---- apply (transformations.py:454-468 -> gaussian.py:1693-1774) -----------------------------------------------
    def rotate(self, R, inv=None, logdet=None, Q=None):
        self.node_X.rotate(R, inv=inv, logdet=logdet)
        if self.plate_axis is not None:
            self.node_X.rotate_plates(Q, plate_axis=self.plate_axis)
        if self.update_alpha:
            self.node_alpha.update()


class RotateGaussianMarkovChain:
    """``RotateGaussianMarkovChain(X, rotA)`` (transformations.py:1112-1452): q(x_n) -> q(R x_n) for a
    ``GaussianMarkovChain`` with unit innovation noise, together with the dynamics A -> R A R^-1 through ``rotA`` (a
    ``RotateGaussianARD`` of the dynamics node, rotated on its variable axis by R^-T and on its plate axis by R).

    With  S0 = <x_0 x_0^T>,  Sn = sum_{n>=1} <x_n x_n^T>,  Sp = sum_{n>=1} <x_{n-1} x_{n-1}^T>,
    Spn = sum_{n>=1} <x_{n-1} x_n^T>,  the initial state N(mu, Lambda^-1) and the dynamics moments <A>, Cov(a_d),
    the part of the bound that changes is
        -1/2 tr(R (Sn + ...) R^T) ... :  yy = tr(R Sn R^T) + tr(Lambda R S0 R^T)
                                         yz = tr(R <A> Spn R^T) + (Lambda mu)^T R <x_0>
                                         zz = tr(R <A> Sp <A>^T R^T) + sum_d r_d^2 tr(Cov(a_d) Sp),   r = column sums of R
        bound_X = -1/2 yy + yz - 1/2 zz + N log|det R|
    (the last term of zz is the same column-sum approximation the plate rotation of A uses)."""

    def __init__(self, X, *args):
        from ...engine.gmc import GaussianMarkovChain
        if not isinstance(X, GaussianMarkovChain):
            raise ValueError("RotateGaussianMarkovChain rotates a GaussianMarkovChain node")
        if len(X.plates) != 0:
            raise NotImplementedError("Rotation of plated chains is not implemented")
        if len(args) == 0:
            raise NotImplementedError()
        if len(args) > 1:
            raise ValueError("Wrong number of arguments")
        self.X_node = X
        self.A_node = X.parents[2]
        self.A_rotator = args[0]
        nu = X.parents[3]
        if getattr(nu, "value", None) is None or not np.all(np.asarray(nu.value) == 1.0):
            raise NotImplementedError("The rotation assumes unit innovation noise")
        self.N = X.N

    def nodes(self):
        return [self.X_node] + self.A_rotator.nodes()

    def rotate(self, R, inv=None, logdet=None):
        if inv is None:
            inv = np.linalg.inv(R)
        if logdet is None:
            logdet = np.linalg.slogdet(R)[1]
        self.X_node.rotate(R, inv=inv, logdet=logdet)
        self.A_rotator.rotate(inv.T, inv=R.T, logdet=-logdet, Q=R)

    def setup(self):
        X, Dm = self.X_node, self.X_node.D
        x, xx, xpxn = X.u
        x0 = D.asarray(x).slice_axis(0, 0, 1)
        self.X0 = _np(x0).reshape(Dm)
        self.X0X0 = _np(D.asarray(xx).slice_axis(0, 0, 1)).reshape(Dm, Dm)
        total = _np(D.sum_product([D.asarray(xx)], [["n", "i", "j"]], ["i", "j"]))
        last = _np(D.asarray(xx).slice_axis(0, self.N - 1, self.N)).reshape(Dm, Dm)
        self.XnXn = total - self.X0X0                               # sum_{n>=1} <x_n x_n^T>
        XpXp = total - last                                        # sum_{n>=1} <x_{n-1} x_{n-1}^T>
        XpXn = _np(D.sum_product([D.asarray(xpxn)], [["n", "i", "j"]], ["i", "j"]))
        u_mu, u_Lam = X.parents[0].get_moments(), X.parents[1].get_moments()
        self.Lambda = np.broadcast_to(_np(u_Lam[0]), (Dm, Dm)).copy()
        self.Lambda_mu_X0 = np.outer(self.Lambda @ np.broadcast_to(_np(u_mu[0]), (Dm,)), self.X0)
        A = np.broadcast_to(_np(self.A_node.u[0]), (Dm, Dm))
        AA = np.broadcast_to(_np(dense(self.A_node.u[1])), (Dm, Dm, Dm))
        if len(self.A_node.plates) != 1:
            raise NotImplementedError("Rotation with time-varying or plated dynamics is not implemented")
        CovA = AA - A[:, :, None] * A[:, None, :]
        self.A_XpXn = A @ XpXn
        self.A_XpXp_A = A @ XpXp @ A.T
        self.CovA_XpXp = np.einsum("dij,ij->d", CovA, XpXp)
        self.A_rotator.setup(plate_axis=-1)

    def _compute_bound(self, R, logdet=None, inv=None, gradient=False, terms=False):
        invR = np.linalg.inv(R) if inv is None else inv
        logdetR = np.linalg.slogdet(R)[1] if logdet is None else logdet
        r = np.sum(R, axis=0)
        R_Sn = R @ self.XnXn
        L_R_S0 = self.Lambda @ R @ self.X0X0
        R_ASA = R @ self.A_XpXp_A
        rc = r * self.CovA_XpXp
        yy = np.sum(R_Sn * R) + np.sum(L_R_S0 * R)
        yz = np.sum((R @ self.A_XpXn) * R) + np.sum(self.Lambda_mu_X0 * R)
        zz = np.sum(R_ASA * R) + np.dot(rc, r)
        value = -0.5 * yy + yz - 0.5 * zz + self.N * logdetR
        if terms:
            value = {self.X_node: value}
        if not gradient:
            return value
        dyy = 2.0 * (R_Sn + L_R_S0)
        dyz = R @ (self.A_XpXn + self.A_XpXn.T) + self.Lambda_mu_X0
        dzz = 2.0 * (R_ASA + rc[None, :])
        return value, -0.5 * dyy + dyz - 0.5 * dzz + self.N * invR.T

    def bound(self, R, logdet=None, inv=None):
        if inv is None:
            inv = np.linalg.inv(R)
        if logdet is None:
            logdet = np.linalg.slogdet(R)[1]
        bX, dX = self._compute_bound(R, logdet=logdet, inv=inv, gradient=True)
        bA, dRA, dQA = self.A_rotator.bound(inv.T, inv=R.T, logdet=-logdet, Q=R)
        # the dynamics' variable axis is rotated by R^-T: d/dR of f(R^-T) = -R^-T (df/dM)^T R^-T
        dRA = -inv.T @ dRA.T @ inv.T
        return bX + bA, dX + dRA + dQA

    def get_bound_terms(self, R, logdet=None, inv=None):
        if inv is None:
            inv = np.linalg.inv(R)
        if logdet is None:
            logdet = np.linalg.slogdet(R)[1]
        out = dict(self.A_rotator.get_bound_terms(inv.T, inv=R.T, logdet=-logdet, Q=R))
        out.update(self._compute_bound(R, logdet=logdet, inv=inv, gradient=False, terms=True))
        return out


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
