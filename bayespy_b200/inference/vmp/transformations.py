"""Rotation parameter expansion: ``RotationOptimizer`` and the rotators of
bayespy/inference/vmp/transformations.py — ``RotateGaussianARD`` (:376-1110, any variable axis of an array with any
plates, with or without a joint rotation of one plate axis, ARD precision with any broadcastable plates, updated or
fixed), ``RotateGaussianMarkovChain`` (:1112-1452, chains over plates, time-varying / per-chain dynamics),
``RotateVaryingMarkovChain`` (:1454-1541), ``RotateSwitchingMarkovChain`` (:1544-1632) and ``RotateMultiple`` (:1635-1677).

Split of work: the plate sums the cost functions need (a handful of D x D [x C x C] arrays per rotated block) are
contracted on the device by the same kernels the VB sweep uses (``bpk_sum_multiply`` on strided views of the moments)
and read back once per ``setup()``; the optimisation over the D x D matrix R (SciPy CG, exactly the reference's call)
runs on the host; applying R to the plated moments is again device work.  The factor-model case of every documented
example (one variable axis, ARD precision over that axis) keeps a two-plate-sum fast path that never materialises the
per-plate covariances of a long node.
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
        if not isinstance(X, GaussianARD) or len(X.dims[0]) < 1:
            raise ValueError("RotateGaussianARD rotates a GaussianARD node with at least one variable axis")
        if subset is not None:
            raise NotImplementedError("Rotation of a subset of the axis is not implemented")
        if len(alpha) > 1:
            raise ValueError("Too many arguments")
        nd = len(X.dims[0])
        if not isinstance(axis, int) or not -nd <= axis < nd:
            raise ValueError("Axis out of bounds")
        self.node_X = X
        self.node_alpha = alpha[0] if alpha else None
        self.update_alpha = bool(alpha)
        if self.update_alpha and not isinstance(self.node_alpha, Gamma):
            raise ValueError("The ARD node must be a Gamma node")
        self.axis = axis % nd - nd                       # counted from the end of the variable axes
        self.D = X.dims[0][self.axis]
        self.plate_axis = None
        # the case every documented factor model has - one variable axis, an ARD precision over exactly that axis (or a
        # fixed one), no plate mixing - keeps its own statistics: two K x K plate sums of the (possibly very long) node
        self._simple = nd == 1 and (not self.update_alpha or tuple(self.node_alpha.plates) == tuple(X.dims[0]))
        self._general = None

    def nodes(self):
        return [self.node_X, self.node_alpha] if self.update_alpha else [self.node_X]

    # ---- statistics (transformations.py:482-640, the branch without plate rotation) -------------------------
    def setup(self, plate_axis=None):
        self.plate_axis = plate_axis
        X, K = self.node_X, self.D
        self._general = None
        if plate_axis is not None and self._simple and len(X.plates) == 1 and plate_axis in (-1, 0):
            return self._setup_with_plate_rotation(plate_axis)
        if plate_axis is not None or not self._simple or not self._simple_parents():
            self._general = _ArrayRotationStatistics(X, self.node_alpha if self.update_alpha else None, self.axis,
                                                     plate_axis)
            return
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

    def _simple_parents(self):
        """True when the prior mean and a fixed precision do not vary over the plates (what the two-plate-sum
        statistics of ``setup`` can express)."""
        X, K = self.node_X, self.D
        u_mu, u_al = X.parents[0].get_moments(), X.parents[1].get_moments()
        if not self.update_alpha:
            a = D.asarray(u_al[0])
            if a.size not in (1, K) or (a.ndim > 1 and any(n != 1 for n in a.shape[:-1])):
                return False
        return True

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
        if self._general is not None:
            return self._general.bound(R, Q, self.node_X, self.node_alpha if self.update_alpha else None,
                                       logdet=logdet, inv=inv, gradient=gradient, terms=terms)
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
        self.node_X.rotate(R, inv=inv, logdet=logdet, axis=self.axis)
        if self.plate_axis is not None:
            self.node_X.rotate_plates(Q, plate_axis=self.plate_axis)
        if self.update_alpha:
            self.node_alpha.update()


class _ArrayRotationStatistics:
    """Rotation of one variable axis (length D, by R) and optionally one plate axis (length C, by Q) of a GaussianARD
    array with plates P and variable shape S, for a prior mean and an ARD precision with any plates that broadcast to
    F = P + S (transformations.py:475-1095 is the reference's treatment of the same problem).

    Index space: F = O x c x v, with v the rotated variable axis, c the rotated plate axis (a dummy axis of length one
    when no plate is rotated) and O every other axis, plate or variable alike.  The precision varies over a sub-grid
    G of O (axes along which it has extent one are summed out) and possibly over c and v.  Plate sums, reduced on the
    device once per ``setup``:

        CovS[g,q,k,l]   = sum_o Cov(x_oqk, x_oql)              MM[g,q,r,k,l] = sum_o <x_oqk> <x_orl>
        Mmu[g,q,p,k,i]  = sum_o <x_oqk> <mu_opi>               mumu[g,p,i]   = sum_o <mu_opi^2>

    After x -> Q x R^T (means mixed exactly over the plate axis, covariances scaled by the squared column sums s_p of Q,
    the approximation GaussianARD.rotate_plates applies) the expected squared deviation from the prior mean is

        e[g,p,i] = s_p^2 (R CovS[g,p] R^T)_ii + sum_qr Q_pq Q_pr (R MM[g,q,r] R^T)_ii - 2 sum_q Q_pq (R Mmu[g,q,p])_ii
                   + mumu[g,p,i]

    summed further over p / i where the precision does not vary.  With a fixed precision the bound changes by
    -1/2 sum <alpha> e; with a Gamma(a0, b0) precision re-optimised after the rotation, b = b0 + e/2 and the bound
    terms are  X: -1/2 (a/b) e - n/2 log b,  alpha: -a0 log b - b0 a/b.  The entropy of q(X) grows by
    |F|/D log|det R| + |F|/C sum_p log|s_p|."""

    def __init__(self, X, alpha, axis, plate_axis):
        P, S = tuple(X.plates), tuple(X.dims[0])
        F = P + S
        nF, nS = len(F), len(S)
        v = nF + axis                                              # position of the rotated variable axis in F
        if plate_axis is None:
            c = None
        else:
            if not isinstance(plate_axis, int):
                raise ValueError("Plate axis must be integer")
            if not -len(P) <= plate_axis < len(P):
                raise ValueError("Plate axis out of bounds")
            c = plate_axis % len(P)
        self.D, self.C = F[v], (F[c] if c is not None else 1)
        self.has_Q = c is not None
        O = [j for j in range(nF) if j != v and j != c]
        self.nF_elems = float(np.prod(F, dtype=np.float64)) if F else 1.0
        # --- moments as device arrays with one key per axis of F
        u_mu, u_al = X.parents[0].get_moments(), X.parents[1].get_moments()
        m = D.asarray(X.u[0]).broadcast_to(F)
        u1 = D.asarray(dense(X.u[1])).broadcast_to(P + S + S).contiguous()
        # second moments between positions that differ along v only: the diagonal over every other variable axis
        st, sh = list(u1.strides), list(u1.shape)
        np_, first, second = len(P), len(P), len(P) + nS
        d_shape = list(F) + [self.D]
        d_strides = list(st[:np_])
        for j in range(nS):
            d_strides.append(st[first + j] if np_ + j == v else st[first + j] + st[second + j])
        d_strides.append(st[second + (v - np_)])
        xxd = DArray(u1.owner, u1.ptr, d_shape, d_strides, u1.dtype)

        def aligned(a):
            a = D.asarray(a)
            if a.ndim > nF:
                a = a.squeeze_leading(nF)
            return a.add_leading(nF - a.ndim)
        mu, mu2 = aligned(u_mu[0]), aligned(u_mu[1])
        if alpha is not None:
            ashape = (1,) * (nF - len(alpha.plates)) + tuple(alpha.plates)
        else:
            ashape = tuple(aligned(u_al[0]).shape)
        for j in range(nF):
            if ashape[j] not in (1, F[j]):
                raise ValueError("The plates of the ARD precision do not broadcast to the rotated array")
        self.av, self.ac = ashape[v], (ashape[c] if c is not None else 1)
        G_axes = [j for j in O if ashape[j] != 1]
        self.G = tuple(F[j] for j in G_axes)
        gk = [("f", j) for j in G_axes]

        def keys(vkey, ckey):
            return [vkey if j == v else (ckey if j == c else ("f", j)) for j in range(nF)]
        sizes = {("f", j): F[j] for j in O}
        cq, cr, cp = (["q"], ["r"], ["p"]) if c is not None else ([], [], [])
        XXS = D.sum_product([xxd], [keys("k", "q") + ["l"]], gk + cq + ["k", "l"], sizes=sizes)
        MMd = D.sum_product([m, m], [keys("k", "q"), keys("l", "q")], gk + cq + ["k", "l"], sizes=sizes)
        MM = D.sum_product([m, m], [keys("k", "q"), keys("l", "r")], gk + cq + cr + ["k", "l"], sizes=sizes)
        szm = dict(sizes)
        szm.update({"i": self.D, "p": self.C})
        Mmu = D.sum_product([m, mu], [keys("k", "q"), keys("i", "p")], gk + cq + cp + ["k", "i"], sizes=szm)
        mumu = D.sum_product([mu2], [keys("i", "p")], gk + cp + ["i"], sizes=szm)
        g = int(np.prod(self.G, dtype=np.int64)) if self.G else 1
        Dn, Cn = self.D, self.C
        self.CovS = (_np(XXS) - _np(MMd)).reshape(g, Cn, Dn, Dn)
        self.MM = _np(MM).reshape(g, Cn, Cn, Dn, Dn)
        self.Mmu = _np(Mmu).reshape(g, Cn, Cn, Dn, Dn)
        self.mumu = _np(mumu).reshape(g, Cn, Dn)
        self.n_o = float(np.prod([F[j] for j in O if j not in G_axes], dtype=np.float64)) if O else 1.0
        # elements of X per element of the precision
        self.n = self.n_o * (Cn if self.ac == 1 else 1) * (Dn if self.av == 1 else 1)
        ash = (g, self.ac, self.av)
        if alpha is not None:
            def to_alpha(a):
                a = np.asarray(a, dtype=np.float64)
                a = a.reshape((1,) * (nF - a.ndim) + a.shape) if a.ndim <= nF else a.reshape(a.shape[a.ndim - nF:])
                full = np.broadcast_to(a, ashape)
                order = G_axes + ([c] if c is not None else []) + [v]
                rest = [j for j in range(nF) if j not in order]
                return np.transpose(full, rest + order).reshape(ash)
            self.a = to_alpha(_np(alpha.phi[1]))
            self.a0 = to_alpha(_np(alpha.parents[0].get_moments()[0]))
            self.b0 = to_alpha(_np(alpha.parents[1].get_moments()[0]))
            self.prec = None
        else:
            a = _np(u_al[0])
            a = a.reshape((1,) * (nF - a.ndim) + a.shape) if a.ndim <= nF else a.reshape(a.shape[a.ndim - nF:])
            order = G_axes + ([c] if c is not None else []) + [v]
            rest = [j for j in range(nF) if j not in order]
            self.prec = np.transpose(a, rest + order).reshape(ash)

    def bound(self, R, Q, node_X, node_alpha, logdet=None, inv=None, gradient=False, terms=False):
        Dn, Cn = self.D, self.C
        if self.has_Q:
            if Q is None:
                raise ValueError("Plates should be rotated but no Q given")
            Q = np.asarray(Q, dtype=np.float64)
        else:
            Q = np.ones((1, 1))
        if logdet is None:
            logdet = np.linalg.slogdet(R)[1]
        if inv is None and gradient:
            inv = np.linalg.inv(R)
        s = np.sum(Q, axis=0)
        cdiag = np.einsum("ik,gpkl,il->gpi", R, self.CovS, R)
        RMR = np.einsum("ik,gqrkl,il->gqri", R, self.MM, R)
        bdiag = np.einsum("pq,pr,gqri->gpi", Q, Q, RMR)
        tdiag = np.einsum("pq,ik,gqpki->gpi", Q, R, self.Mmu)
        e_full = (s * s)[None, :, None] * cdiag + bdiag - 2.0 * tdiag + self.mumu
        e = e_full
        if self.ac == 1:
            e = e.sum(axis=1, keepdims=True)
        if self.av == 1:
            e = e.sum(axis=2, keepdims=True)
        n = self.n
        if node_alpha is not None:
            rate = self.b0 + 0.5 * e
            prec = self.a / rate
            logprec = -np.log(rate)
            hyper = np.sum(self.a0 * logprec) - np.sum(self.b0 * prec)
        else:
            prec, logprec, hyper = self.prec, 0.0, 0.0
        entropy_gain = self.nF_elems / Dn * logdet + self.nF_elems / Cn * np.sum(np.log(np.abs(s)))
        fit = -0.5 * np.sum(prec * e) + 0.5 * n * np.sum(logprec)
        if terms:
            out = {node_X: fit + entropy_gain}
            if node_alpha is not None:
                out[node_alpha] = hyper
            return out
        value = fit + hyper + entropy_gain
        if not gradient:
            return value
        if node_alpha is not None:
            w = -0.5 * prec + 0.5 * (0.5 * e * prec - 0.5 * n - self.a0 + self.b0 * prec) / rate
        else:
            w = -0.5 * prec
        W = np.broadcast_to(w, e_full.shape)
        B = np.einsum("pq,pr,gqrkl->gpkl", Q, Q, self.MM)
        T = np.einsum("pq,gqpki->gpki", Q, self.Mmu)
        dR = 2.0 * np.einsum("gpi,p,ik,gpkl->il", W, s * s, R, self.CovS) \
            + np.einsum("gpi,ik,gpkl->il", W, R, B + np.swapaxes(B, -1, -2)) \
            - 2.0 * np.einsum("gpi,gpli->il", W, T) + self.nF_elems / Dn * inv.T
        if not self.has_Q:
            return value, dR
        dQ = 2.0 * np.einsum("gpi,pr,gqri->pq", W, Q, RMR) - 2.0 * np.einsum("gpi,ik,gqpki->pq", W, R, self.Mmu)
        col = 2.0 * s * np.einsum("gpi,gpi->p", W, cdiag) + self.nF_elems / Cn / s
        return value, dR, dQ + col[None, :]


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
        if len(args) == 0:
            raise NotImplementedError()
        if len(args) > 1:
            raise ValueError("Wrong number of arguments")
        if len(X.parents) > 4:
            raise NotImplementedError("Rotation of a chain with input signals is not implemented")
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

    def _dynamics_statistics(self, xx_head, xpxn, pk, sizes):
        """Sums over chains and time steps that couple the dynamics with the chain's second moments:
        sum A_n <x_{n-1} x_n^T>,  sum A_n <x_{n-1} x_{n-1}^T> A_n^T  and, per state d, sum tr(Cov(a_nd) <x_{n-1} x_{n-1}^T>).
        The dynamics may have a time axis of length N-1 or 1 and fewer plates than the chain."""
        X = self.X_node
        dist = X._distribution
        u_par = X.moments_from_parents()
        _, _, _, _, A, AA, _, _ = dist._parents(u_par[0], u_par[1], u_par[2], u_par[3])
        npl = len(pk)
        ka = pk[npl - (A.ndim - 3):]
        A_XpXn = D.sum_product([A, xpxn], [ka + ["n", "i", "k"], pk + ["n", "k", "j"]], ["i", "j"], sizes=sizes)
        A_XpXp_A = D.sum_product([A, xx_head, A], [ka + ["n", "i", "k"], pk + ["n", "k", "l"], ka + ["n", "j", "l"]],
                                 ["i", "j"], sizes=sizes)
        aa = D.sum_product([AA, xx_head], [ka + ["n", "d", "k", "l"], pk + ["n", "k", "l"]], ["d"], sizes=sizes)
        mm = D.sum_product([A, xx_head, A], [ka + ["n", "d", "k"], pk + ["n", "k", "l"], ka + ["n", "d", "l"]], ["d"],
                           sizes=sizes)
        return _np(A_XpXn), _np(A_XpXp_A), _np(aa) - _np(mm)

    def setup(self):
        X, Dm, N = self.X_node, self.X_node.D, self.N
        P = tuple(X.plates)
        npl = len(P)
        pk = [("p", j) for j in range(npl, 0, -1)]
        psizes = {k: n for k, n in zip(pk, P)}          # extents of the chain plates (parents may lack some of them)
        sizes = dict(psizes)
        sizes["n"] = N - 1
        x, xx, xpxn = [D.asarray(v) for v in X.u]
        x = x.broadcast_to(P + (N, Dm))
        xx = xx.broadcast_to(P + (N, Dm, Dm))
        xpxn = xpxn.broadcast_to(P + (N - 1, Dm, Dm))
        x0 = x.slice_axis(npl, 0, 1).reshape(P + (Dm,))
        x0x0 = xx.slice_axis(npl, 0, 1).reshape(P + (Dm, Dm))
        xx_tail = xx.slice_axis(npl, 1, N)
        xx_head = xx.slice_axis(npl, 0, N - 1)
        self.XnXn = _np(D.sum_product([xx_tail], [pk + ["n", "i", "j"]], ["i", "j"]))     # sum_{n>=1} <x_n x_n^T>
        u_mu, u_Lam = X.parents[0].get_moments(), X.parents[1].get_moments()
        Lam, mu = D.asarray(u_Lam[0]), D.asarray(u_mu[0])
        kL, km = pk[npl - (Lam.ndim - 2):], pk[npl - (mu.ndim - 1):]
        # sum over chains of Lambda (x) <x_0 x_0^T> and of (Lambda mu) <x_0>^T
        self.L_X0X0 = _np(D.sum_product([Lam, x0x0], [kL + ["i", "j"], pk + ["k", "l"]], ["i", "j", "k", "l"], sizes=psizes))
        self.Lambda_mu_X0 = _np(D.sum_product([Lam, mu, x0], [kL + ["i", "k"], km + ["k"], pk + ["j"]], ["i", "j"],
                                              sizes=psizes))
        self.A_XpXn, self.A_XpXp_A, self.CovA_XpXp = self._dynamics_statistics(xx_head, xpxn, pk, sizes)
        self.n_chains = float(np.prod(P, dtype=np.float64)) if P else 1.0
        self.A_rotator.setup(plate_axis=-1)

    def _compute_bound(self, R, logdet=None, inv=None, gradient=False, terms=False):
        invR = np.linalg.inv(R) if inv is None else inv
        logdetR = np.linalg.slogdet(R)[1] if logdet is None else logdet
        r = np.sum(R, axis=0)
        R_Sn = R @ self.XnXn
        R_ASA = R @ self.A_XpXp_A
        rc = r * self.CovA_XpXp
        T4 = self.L_X0X0                                             # T4[i,j,k,l] = sum_b Lambda_b[i,j] <x_0 x_0^T>_b[k,l]
        init = np.einsum("ijkl,jk,il->", T4, R, R)                   # sum_b tr(Lambda_b R <x_0 x_0^T>_b R^T)
        yy = np.sum(R_Sn * R) + init
        yz = np.sum((R @ self.A_XpXn) * R) + np.sum(self.Lambda_mu_X0 * R)
        zz = np.sum(R_ASA * R) + np.dot(rc, r)
        entropy = self.N * self.n_chains
        value = -0.5 * yy + yz - 0.5 * zz + entropy * logdetR
        if terms:
            value = {self.X_node: value}
        if not gradient:
            return value
        dinit = np.einsum("iabl,il->ab", T4, R) + np.einsum("ajkb,jk->ab", T4, R)
        dyy = 2.0 * R_Sn + dinit
        dyz = R @ (self.A_XpXn + self.A_XpXn.T) + self.Lambda_mu_X0
        dzz = 2.0 * (R_ASA + rc[None, :])
        return value, -0.5 * dyy + dyz - 0.5 * dzz + entropy * invR.T

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


class RotateVaryingMarkovChain(RotateGaussianMarkovChain):
    """``RotateVaryingMarkovChain(X, B, S, B_rotator)`` (transformations.py:1454-1541): the chain's dynamics are
    A_n = sum_k s_nk B_k (``SumMultiply('dk,k->d', B, S)`` under a ``GaussianMarkovChain``, which is also what
    ``VaryingGaussianMarkovChain`` builds).  q(x_n) -> q(R x_n) goes with B_k -> R B_k R^-1 through ``B_rotator`` (a
    ``RotateGaussianARD`` of B on its first variable axis, rotated by R^-T, and on its last plate axis, by R); the
    weights S are left alone.

    Only the statistics that couple the dynamics with the chain differ from the parent class: with
    <s s^T>_n the second moment of the weights and Sp_n = <x_{n-1} x_{n-1}^T>,
        G[e,f] = sum_n sum_kk' <s_k s_k'>_n  mean(b_ek)^T Sp_n mean(b_fk')       (transforms exactly, as R G R^T)
        c[e]   = sum_n sum_kk' <s_k s_k'>_n  tr(Cov(b_ek, b_ek') Sp_n)            (scaled by the squared column sums of R)"""

    def __init__(self, X, B, S, B_rotator):
        super().__init__(X, B_rotator)
        from ...engine.dot import SumMultiply
        A = self.A_node
        ok = isinstance(A, SumMultiply) and len(A.parents) == 2 and len(A.in_keys[0]) == 2 and \
            list(A.in_keys[1]) == [A.in_keys[0][1]] and list(A.out_keys) == [A.in_keys[0][0]]
        if not ok:
            raise ValueError("The dynamics of the chain must be SumMultiply('dk,k->d', B, S)")
        self.B_node, self.S_node = B, S

    def _dynamics_statistics(self, xx_head, xpxn, pk, sizes):
        X, A = self.X_node, self.A_node
        u_par = X.moments_from_parents()
        _, _, _, _, Abar, _, _, _ = X._distribution._parents(u_par[0], u_par[1], u_par[2], u_par[3])
        npl = len(pk)
        ka = pk[npl - (Abar.ndim - 3):]
        A_XpXn = D.sum_product([Abar, xpxn], [ka + ["n", "i", "k"], pk + ["n", "k", "j"]], ["i", "j"], sizes=sizes)
        u_B, u_S = A.parents[0].get_moments(), A.parents[1].get_moments()
        b, bb = D.asarray(u_B[0]), D.asarray(dense(u_B[1]))          # plates (.., T|1, D) + (D, K) [+ (D, K)]
        sv, ss = D.asarray(u_S[0]), D.asarray(dense(u_S[1]))         # plates (.., T|1, 1) + (K,) [+ (K,)]
        plate_keys = pk + ["n", "e"]

        def pl(a, nd, row="e"):
            ks = list(plate_keys[len(plate_keys) - (a.ndim - nd):]) if a.ndim > nd else []
            return [row if k == "e" else k for k in ks]
        kb, kb2 = pl(b, 2), pl(b, 2, row="f")
        kS = pl(ss, 2, row="z")
        Sp = [pk + ["n", "i", "j"]]
        G = D.sum_product([b, ss, xx_head, b], [kb + ["i", "k"], kS + ["k", "m"]] + Sp + [kb2 + ["j", "m"]], ["e", "f"],
                          sizes=sizes)
        tot = D.sum_product([bb, ss, xx_head], [pl(bb, 4) + ["i", "k", "j", "m"], kS + ["k", "m"]] + Sp, ["e"], sizes=sizes)
        G = _np(G)
        Dm = X.D
        G = np.broadcast_to(G, (Dm, Dm)) if G.shape != (Dm, Dm) else G
        return _np(A_XpXn), G, np.broadcast_to(_np(tot), (Dm,)) - np.diag(G)


class RotateSwitchingMarkovChain(RotateGaussianMarkovChain):
    """``RotateSwitchingMarkovChain(X, B, Z, B_rotator)`` (transformations.py:1544-1632): at every step the dynamics are
    one of K matrices, A_n = B_{z_n} (``Gate(Z, B, gated_plate=-2)`` under a ``GaussianMarkovChain``, which is what
    ``SwitchingGaussianMarkovChain`` builds); B has plates (..., K, D).  q(x_n) -> q(R x_n) goes with B_k -> R B_k R^-1
    through ``B_rotator`` (``RotateGaussianARD(B)``: variable axis by R^-T, last plate axis by R).  With the selection
    probabilities p_nk = <z_nk>:
        G[e,f] = sum_n sum_k p_nk mean(b_ke)^T Sp_n mean(b_kf),      c[e] = sum_n sum_k p_nk tr(Cov(b_ke) Sp_n)."""

    def __init__(self, X, B, Z, B_rotator):
        super().__init__(X, B_rotator)
        from ...engine.gate import Gate
        A = self.A_node
        if not (isinstance(A, Gate) and A.gated_plate == -2):
            raise ValueError("The dynamics of the chain must be Gate(Z, B, gated_plate=-2)")
        K, Dm = A.K, X.D
        if tuple(B.plates[-2:]) != (K, Dm):
            raise ValueError("Incorrect plates in B")
        if len(B.dims[0]) != 1:
            raise ValueError("B should have exactly one variable axis")
        self.B_node, self.Z_node = B, Z

    def _dynamics_statistics(self, xx_head, xpxn, pk, sizes):
        X, A = self.X_node, self.A_node
        u_par = X.moments_from_parents()
        _, _, _, _, Abar, _, _, _ = X._distribution._parents(u_par[0], u_par[1], u_par[2], u_par[3])
        npl = len(pk)
        ka = pk[npl - (Abar.ndim - 3):]
        A_XpXn = D.sum_product([Abar, xpxn], [ka + ["n", "i", "k"], pk + ["n", "k", "j"]], ["i", "j"], sizes=sizes)
        pz = D.asarray(A.parents[0].get_moments()[0])                 # plates (.., N-1 | 1, 1) + (K,)
        u_B = A.parents[1].get_moments()
        b, bb = D.asarray(u_B[0]), D.asarray(dense(u_B[1]))           # plates (.., K, D) + (D,) [+ (D,)]
        zkeys = (pk + ["n", "z"])
        kz = list(zkeys[len(zkeys) - (pz.ndim - 1):]) + ["c"]

        def kB(a, nd, row):
            ks = pk + ["c", row]
            return list(ks[len(ks) - (a.ndim - nd):])
        Sp = [pk + ["n", "i", "j"]]
        G = D.sum_product([pz, b, xx_head, b], [kz, kB(b, 1, "e") + ["i"]] + Sp + [kB(b, 1, "f") + ["j"]], ["e", "f"],
                          sizes=sizes)
        tot = D.sum_product([pz, bb, xx_head], [kz, kB(bb, 2, "e") + ["i", "j"]] + Sp, ["e"], sizes=sizes)
        G = _np(G)
        return _np(A_XpXn), G, _np(tot) - np.diag(G)


class RotateMultiple:
    """The same rotation applied to several blocks at once; their cost functions add (transformations.py:1635-1677)."""

    def __init__(self, *rotators):
        self.rotators = rotators

    def nodes(self):
        return [n for r in self.rotators for n in r.nodes()]

    def rotate(self, R, inv=None, logdet=None):
        for r in self.rotators:
            r.rotate(R, inv=inv, logdet=logdet)

    def setup(self):
        for r in self.rotators:
            r.setup()

    def bound(self, R, logdet=None, inv=None):
        value, grad = 0.0, 0.0
        for r in self.rotators:
            b, db = r.bound(R, logdet=logdet, inv=inv)
            value, grad = value + b, grad + db
        return value, grad

    def get_bound_terms(self, *args, **kwargs):
        out = {}
        for r in self.rotators:
            out.update(r.get_bound_terms(*args, **kwargs))
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
