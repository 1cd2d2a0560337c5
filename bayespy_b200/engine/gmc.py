"""Gaussian Markov chain node on device (nodes/gaussian_markov_chain.py:270-927):

    x_0 ~ N(mu, Lambda^-1),   x_n ~ N(A_{n-1} [x_{n-1}; z_{n-1}], diag(nu_{n-1})^-1),  n = 1..N-1

with optional input signals z (a Gaussian-like parent with plates (..., N-1 | 1) and K-dimensional values; the rows of A
then have length D + K), dynamics that may carry a gamma scale of their own (A a Gaussian-gamma node),
time-invariant (A plates (..., 1, D) or (D,)) or per-step (plates (..., N-1, D)) dynamics and independent
chains over leading plates, in the reference's plate layout.

Moments u = [<x_n> (N,D), <x_n x_n^T> (N,D,D), <x_n x_{n+1}^T> (N-1,D,D)]; the time axis is part of the
variable dims exactly as in the reference.  The natural parameters form a block-tridiagonal precision
(:542-627); the update is one call of ``bpk_block_banded_solve`` (linalg.py:468-575, the RTS smoother in
information form) instead of the reference's Python loop over time steps.

The reference routes (mu, Lambda) and (A, nu) through the joint-moment wrapper nodes
WrapToGaussianWishart / WrapToGaussianGamma (gaussian.py:2299-2527); here the node has four plain parents
and forms those products itself, so the messages arriving at mu, Lambda, A and nu are the same arrays.
"""
import numpy as np

from .. import _bpk
from .. import darray as D
from ..darray import DArray
from .expfam import Distribution, ExponentialFamily
from .gaussian import LOG2PI, dense, ensure_gamma, ensure_gaussian
from .node import Deterministic, Node
from .wishart import ensure_wishart


class _SingleChainDistribution(Distribution):
    """The plates == () case, kept verbatim as measured and validated on the GPU in round 1; the general
    class below delegates to it whenever the chain has no plates."""


    def __init__(self, N, Dm):
        self.N, self.D = int(N), int(Dm)

    # -- plates: (mu, Lambda) see the chain's plates; A and nu additionally carry the state axis (D,)
    def plates_to_parent(self, index, plates):
        return tuple(plates) if index < 2 else tuple(plates) + (self.D,)

    def plates_from_parent(self, index, plates):
        return tuple(plates) if index < 2 else tuple(plates[:-1])

    def compute_weights_to_parent(self, index, weights):
        w = np.asarray(weights)
        return w if index < 2 else w.reshape(w.shape + (1,))

    def _check_static(self, u_A, u_nu):
        if u_A[0].ndim != 2 or D.asarray(u_nu[0]).ndim > 1:
            raise NotImplementedError("GaussianMarkovChain: plated chains and time-varying dynamics are not "
                                      "supported yet (A must have plates (D,), nu plates (D,))")

    # -- natural parameters (gaussian_markov_chain.py:542-627)
    def compute_phi_from_parents(self, u_mu, u_Lambda, u_A, u_nu, mask=True):
        N, Dm = self.N, self.D
        self._check_static(u_A, u_nu)
        Lam = D.asarray(u_Lambda[0]).reshape((Dm, Dm))
        mu = D.asarray(u_mu[0]).reshape((Dm,))
        A = u_A[0]                                   # (D, D): row d = <a_d>
        AA = dense(u_A[1])                           # (D, D, D): <a_d a_d^T>
        nu = D.asarray(u_nu[0]).broadcast_to((Dm,)).contiguous()
        phi0 = DArray.zeros((N, Dm))
        D.sum_product([Lam, mu], [["i", "j"], ["j"]], ["i"], out=phi0.slice_axis(0, 0, 1).reshape((Dm,)))
        phi1 = DArray.zeros((N, Dm, Dm))
        D._ew("AFFINE", (Dm, Dm), phi1.slice_axis(0, 0, 1).reshape((Dm, Dm)), [Lam], alpha=-0.5, beta=0.0)
        if N > 1:
            # diagonal blocks n >= 1: -1/2 diag(nu);  blocks n <= N-2: -1/2 sum_d nu_d <a_d a_d^T>
            tail = phi1.slice_axis(0, 1, N)
            D._ew("AFFINE", (N - 1, Dm), tail.diag_view(1), [nu.reshape((1, Dm))], alpha=-0.5, beta=0.0)
            S = D.sum_product([nu, AA], [["d"], ["d", "i", "j"]], ["i", "j"], scale=-0.5)
            head = phi1.slice_axis(0, 0, N - 1)
            D._ew("ADD", (N - 1, Dm, Dm), head, [head, S.reshape((1, Dm, Dm))])
        # super-diagonal blocks (sum of super and sub): phi2[n, i, j] = nu_j <A>[j, i]
        phi2 = DArray.empty((max(N - 1, 0), Dm, Dm))
        if N > 1:
            nuA_T = D.mul(A, nu.reshape((Dm, 1))).swap_last2()
            D.copy_into(phi2, nuA_T.reshape((1, Dm, Dm)))
        return [phi0, phi1, phi2]

    # -- E[log normaliser of the prior] (:251-267, :629-657)
    def compute_cgf_from_parents(self, u_mu, u_Lambda, u_A, u_nu):
        Dm = self.D
        Lam = D.asarray(u_Lambda[0]).reshape((Dm, Dm))
        mumu = dense(u_mu[1]).reshape((Dm, Dm))
        t = D.sum_product([Lam, mumu], [["i", "j"], ["i", "j"]], [])
        g = D.axpby(-0.5, t, 0.5, D.asarray(u_Lambda[1]).reshape(()))
        lognu = D.asarray(u_nu[1]).broadcast_to((Dm,))
        s = D.sum_product([lognu], [["d"]], [], scale=0.5 * (self.N - 1))
        return D.add(g, s)

    # -- smoother (:89-123)
    def compute_moments_and_cgf(self, phi, mask=True):
        N, Dm = self.N, self.D
        be = _bpk.get()
        phi = [D.asarray(v) for v in phi]
        y = phi[0].contiguous()
        A = D.mul(phi[1], -2.0)
        B = D.mul(phi[2], -1.0) if N > 1 else DArray.empty((1,))
        V, x = DArray.empty((N, Dm, Dm)), DArray.empty((N, Dm))
        C = DArray.empty((max(N - 1, 1), Dm, Dm))
        ld = DArray.empty(())
        be.block_banded_solve(A.ptr, B.ptr, y.ptr, 1, N, Dm, V.ptr, C.ptr, x.ptr, ld.ptr, True)
        u1 = D.add(V, D.mul(x.add_trailing(1), x.expand_dims(-2)))
        if N > 1:
            xp = x.slice_axis(0, 0, N - 1)
            xn = x.slice_axis(0, 1, N)
            u2 = D.add(C.slice_axis(0, 0, N - 1), D.mul(xp.add_trailing(1), xn.expand_dims(-2)))
        else:
            u2 = DArray.empty((0, Dm, Dm))
        g = D.axpby(-0.5, D.sum_product([x, y], [["n", "i"], ["n", "i"]], []), 0.5, ld)
        return [x, u1, u2], g

    def compute_fixed_moments_and_f(self, x, mask=True):
        x = np.asarray(x, dtype=np.float64)
        if x.shape[-2:] != (self.N, self.D):
            raise ValueError("Invalid shape")
        u1 = x[..., :, None] * x[..., None, :]
        u2 = x[..., :-1, :, None] * x[..., 1:, None, :]
        return [D.asarray(x), D.asarray(u1), D.asarray(u2)], -0.5 * self.N * self.D * LOG2PI

    # -- messages (:443-527 combined with gaussian.py:2351-2371 / :2496-2522)
    def compute_message_to_parent(self, parent, index, u, u_mu, u_Lambda, u_A, u_nu):
        N, Dm = self.N, self.D
        x, xx, xpxn = u
        if index in (0, 1):
            x0 = x.slice_axis(0, 0, 1).reshape((Dm,))
            x0x0 = xx.slice_axis(0, 0, 1).reshape((Dm, Dm))
            if index == 0:
                Lam = D.asarray(u_Lambda[0]).reshape((Dm, Dm))
                return [D.sum_product([Lam, x0], [["i", "j"], ["j"]], ["i"]), D.mul(Lam, -0.5)]
            mu = D.asarray(u_mu[0]).reshape((Dm,))
            mumu = dense(u_mu[1]).reshape((Dm, Dm))
            xm = D.mul(x0.reshape((Dm, 1)), mu.reshape((1, Dm)))
            t = D.add(D.sub(D.sub(x0x0, xm), xm.swap_last2()), mumu)
            return [D.mul(t, -0.5), D.asarray(0.5)]
        self._check_static(u_A if u_A is not None else [DArray.empty((Dm, Dm))], u_nu if u_nu is not None else [0.0])
        if N < 2:
            return [None, None]
        # time sums of the chain's second moments
        Sxx_head = D.sum_product([xx.slice_axis(0, 0, N - 1)], [["n", "i", "j"]], ["i", "j"])     # sum_{n<N-1} <x_n x_n^T>
        Sxpxn = D.sum_product([xpxn], [["n", "i", "j"]], ["i", "j"])                              # sum_n <x_n x_{n+1}^T>
        if index == 2:
            nu = D.asarray(u_nu[0]).broadcast_to((Dm,))
            # to a_d: [nu_d sum_n <x_{n+1,d} x_n>, -1/2 nu_d sum_n <x_n x_n^T>]
            m0 = D.mul(Sxpxn.swap_last2(), nu.reshape((Dm, 1)))
            m1 = D.mul(D.mul(Sxx_head.reshape((1, Dm, Dm)), nu.reshape((Dm, 1, 1))), -0.5)
            return [m0, m1]
        if index == 3:
            A = u_A[0]
            AA = dense(u_A[1])
            dxx = D.sum_product([xx.slice_axis(0, 1, N).diag_view(1)], [["n", "d"]], ["d"], scale=-0.5)
            a = D.sum_product([Sxpxn, A], [["i", "d"], ["d", "i"]], ["d"])
            b = D.sum_product([Sxx_head, AA], [["i", "j"], ["d", "i", "j"]], ["d"], scale=-0.5)
            return [D.add(D.add(dxx, a), b), DArray.full((Dm,), 0.5 * (N - 1))]
        raise ValueError("Parent index out of bounds")



def _pk(n):
    """Plate keys, right-aligned: n keys ('p', n) ... ('p', 1)."""
    return [("p", j) for j in range(n, 0, -1)]


class GaussianMarkovChainDistribution(Distribution):
    """General case, in the reference's plate layout (gaussian_markov_chain.py:660-720): with chain plates P,
    (mu, Lambda) have plates P and (A, nu) have plates P + (N-1 | 1, D) — the second-last plate of the dynamics
    is the TIME axis (1 = time-invariant, N-1 = one transition matrix / innovation precision per step), the last
    one the state axis.  phi / u are P+(N,D), P+(N,D,D), P+(N-1,D,D).  Parents may have fewer (broadcast) plates.

    ``TA`` / ``Tn``: time extents (1 or N-1) of the A and nu parents.  When both are 1 the messages to A and nu
    are summed over time here and carry a time axis of length 1 (so that no (N-1, D, D, D) array is ever formed);
    otherwise they keep the time axis and the generic plate reduction (node.py:570-655) does what is left."""

    def __init__(self, N, Dm, plates=(), TA=1, Tn=1, plain=True, K=0):
        self.N, self.D = int(N), int(Dm)
        self.plates = tuple(int(p) for p in plates)
        self.TA, self.Tn = int(TA), int(Tn)
        self.static = self.TA == 1 and self.Tn == 1
        # input signals (:276, :485-540): x_n = A [x_{n-1}; z_{n-1}] + noise with rows of A of length E = D + K
        self.K = int(K)
        self.E = self.D + self.K
        # ``plain``: the dynamics parents have exactly the plates (D,) (no explicit time axis of length one either)
        self._single = _SingleChainDistribution(N, Dm) if (not self.plates and self.static and plain and not K) else None

    def _t_msg(self):
        return 1 if self.static else self.N - 1

    # -- plates (:660-706)
    def plates_to_parent(self, index, plates):
        if self._single is not None:
            return self._single.plates_to_parent(index, plates)
        if index == 4:
            return tuple(plates) + (self.N - 1,)
        return tuple(plates) if index < 2 else tuple(plates) + (self._t_msg(), self.D)

    def plates_from_parent(self, index, plates):
        if index < 2:
            return tuple(plates)
        if index == 4:
            return tuple(plates[:-1])
        return tuple(plates[:-2])

    def compute_weights_to_parent(self, index, weights):
        if self._single is not None:
            return self._single.compute_weights_to_parent(index, weights)
        w = np.asarray(weights)
        if index == 4:
            return w.reshape(w.shape + (1,))
        return w if index < 2 else w.reshape(w.shape + (1, 1))

    def _parents(self, u_mu, u_Lambda, u_A, u_nu):
        """Parent moments with their own (possibly shorter) plate axes in front of fixed trailing axes:
        A: PA+(TA,D,D) (row d = <a_d>), AA: PA+(TA,D,D,D), nu / lognu: Pn+(Tn,D)."""
        Dm = self.D
        mu = D.asarray(u_mu[0]) if u_mu is not None else None                 # Pm + (D,)
        mumu = dense(u_mu[1]) if u_mu is not None else None                  # Pm + (D, D)
        Lam = D.asarray(u_Lambda[0]) if u_Lambda is not None else None        # PL + (D, D)
        logdetL = D.asarray(u_Lambda[1]) if u_Lambda is not None else None    # PL
        A = AA = None
        if u_A is not None:
            A = D.asarray(u_A[0])
            AA = dense(u_A[1])
            if A.ndim == 1:
                A, AA = A.add_leading(1), AA.add_leading(1)
            # the state plate may be stored compressed (a prior shared by all rows)
            A = A.broadcast_to(tuple(A.shape[:-2]) + (Dm, self.E))
            AA = AA.broadcast_to(tuple(AA.shape[:-3]) + (Dm, self.E, self.E))
            if A.ndim == 2:                                                   # plates (D,): no time axis
                A, AA = A.add_leading(1), AA.add_leading(1)
            if A.shape[-3] not in (1, self.N - 1):
                raise ValueError("The second last plate of the dynamics matrix should have length one or N-1")
        nu = lognu = None
        if u_nu is not None:
            nu, lognu = D.asarray(u_nu[0]), D.asarray(u_nu[1])
            if nu.ndim == 0:
                nu, lognu = nu.reshape((1, 1)), lognu.reshape((1, 1))
            elif nu.ndim == 1:
                nu, lognu = nu.reshape((1, nu.shape[0])), lognu.reshape((1, lognu.shape[0]))
            if nu.shape[-1] != Dm:
                nu = nu.broadcast_to(tuple(nu.shape[:-1]) + (Dm,))
                lognu = lognu.broadcast_to(tuple(lognu.shape[:-1]) + (Dm,))
            if nu.shape[-2] not in (1, self.N - 1):
                raise ValueError("The second last plate of the innovation precision should have length one or N-1")
        return mu, mumu, Lam, logdetL, A, AA, nu, lognu

    def _dynamics_scale(self, u_A):
        """(<tau>, <log tau>) of a Gaussian-gamma dynamics parent, plates PA + (TA, D) like nu; None for a Gaussian one.
        With it, u_A[0] = <tau a>, u_A[1] = <tau a a^T> and the innovation precision of row d is nu_d tau_d
        (WrapToGaussianGamma, gaussian.py:2339-2348)."""
        if u_A is None or len(u_A) < 4:
            return None
        t, lt = D.asarray(u_A[2]), D.asarray(u_A[3])
        if t.ndim == 0:
            t, lt = t.reshape((1, 1)), lt.reshape((1, 1))
        elif t.ndim == 1:
            t, lt = t.reshape((1, t.shape[0])), lt.reshape((1, lt.shape[0]))
        return t, lt

    def _inputs(self, u_inputs):
        """z: Pz + (Tz, K), zz: Pz + (Tz, K, K) of the input-signal parent (None without inputs)."""
        if not self.K:
            return None, None
        z, zz = D.asarray(u_inputs[0][0]), dense(u_inputs[0][1])
        if z.ndim == 1:
            z, zz = z.add_leading(1), zz.add_leading(1)
        return z, zz

    def _regressor_moments(self, u, z, zz):
        """<c_n x_{n+1}^T> (P+(N-1, E, D)) and <c_n c_n^T> (P+(N-1, E, E)) of the regressor c_n = [x_n; z_n]: the chain's
        own moments when there are no inputs, else the blocks of :485-500 placed side by side."""
        N, Dm, P, E = self.N, self.D, self.plates, self.E
        npl = len(P)
        x, xx, xpxn = u
        xx_head = xx.slice_axis(npl, 0, N - 1)
        if not self.K:
            return xpxn, xx_head
        xh = x.slice_axis(npl, 0, N - 1)                                  # x_n,     P + (N-1, D)
        xn = x.slice_axis(npl, 1, N)                                      # x_{n+1}
        cxn = DArray.zeros(P + (N - 1, E, Dm))
        D.copy_into(cxn.slice_axis(npl + 1, 0, Dm), xpxn)
        D.copy_into(cxn.slice_axis(npl + 1, Dm, E), D.mul(z.add_trailing(1), xn.expand_dims(-2)))
        cc = DArray.zeros(P + (N - 1, E, E))
        top = cc.slice_axis(npl + 1, 0, Dm)
        bot = cc.slice_axis(npl + 1, Dm, E)
        xz = D.mul(xh.add_trailing(1), z.expand_dims(-2))                 # x_n z_n^T,  (.., N-1, D, K)
        D.copy_into(top.slice_axis(npl + 2, 0, Dm), xx_head)
        D.copy_into(top.slice_axis(npl + 2, Dm, E), xz)
        D.copy_into(bot.slice_axis(npl + 2, 0, Dm), xz.swap_last2())
        D.copy_into(bot.slice_axis(npl + 2, Dm, E), zz)
        return cxn, cc

    # -- natural parameters (gaussian_markov_chain.py:542-627)
    def compute_phi_from_parents(self, u_mu, u_Lambda, u_A, u_nu, *u_inputs, mask=True):
        if self._single is not None and len(u_A) < 4:
            return self._single.compute_phi_from_parents(u_mu, u_Lambda, u_A, u_nu, mask=mask)
        N, Dm, P = self.N, self.D, self.plates
        npl = len(P)
        mu, _, Lam, _, A_full, AA_full, nu, _ = self._parents(u_mu, u_Lambda, u_A, u_nu)
        # the state block of the dynamics
        A = A_full.slice_axis(A_full.ndim - 1, 0, Dm)
        AA = AA_full.slice_axis(AA_full.ndim - 2, 0, Dm).slice_axis(AA_full.ndim - 1, 0, Dm)
        scale = self._dynamics_scale(u_A)
        nu_diag = nu if scale is None else D.mul(nu, scale[0])
        pk = _pk(npl)
        # x_0: phi0[..., 0, :] = <Lambda> <mu>,  phi1[..., 0, :, :] = -1/2 <Lambda>
        Lmu = D.sum_product([Lam, mu], [pk[npl - (Lam.ndim - 2):] + ["i", "j"], pk[npl - (mu.ndim - 1):] + ["j"]], pk + ["i"])
        phi0 = DArray.zeros(P + (N, Dm))
        D.copy_into(phi0.slice_axis(npl, 0, 1), Lmu.reshape(tuple(Lmu.shape[:-1]) + (1, Dm)))
        phi1 = DArray.zeros(P + (N, Dm, Dm))
        first = phi1.slice_axis(npl, 0, 1)
        D._ew("AFFINE", first.shape, first, [Lam.reshape(tuple(Lam.shape[:-2]) + (1, Dm, Dm))], alpha=-0.5, beta=0.0)
        phi2 = DArray.empty(P + (max(N - 1, 0), Dm, Dm))
        if N > 1:
            # blocks n >= 1: -1/2 diag(nu_{n-1});  blocks n <= N-2: -1/2 sum_d nu_{n,d} <a_{n,d} a_{n,d}^T>
            tail = phi1.slice_axis(npl, 1, N)
            D._ew("AFFINE", tail.diag_view(1).shape, tail.diag_view(1), [nu_diag], alpha=-0.5, beta=0.0)
            nk, ak = nu.ndim - 2, AA.ndim - 4
            S = D.sum_product([nu, AA], [pk[npl - nk:] + ["n", "d"], pk[npl - ak:] + ["n", "d", "i", "j"]],
                              pk + ["n", "i", "j"], scale=-0.5)
            head = phi1.slice_axis(npl, 0, N - 1)
            D._ew("ADD", head.shape, head, [head, S])
            # super-diagonal blocks (sum of super and sub): phi2[..., n, i, j] = nu_{n,j} <A_n>[j, i]
            nuA_T = D.mul(A, nu.add_trailing(1)).swap_last2()
            D.copy_into(phi2, nuA_T)
            if self.K:
                # effect of the input signals (:608-616): phi0[n+1] += nu B z_n,  phi0[n] -= sum_d nu_d <a_d b_d^T> z_n
                z, _ = self._inputs(u_inputs)
                E = self.E
                B = A_full.slice_axis(A_full.ndim - 1, Dm, E)
                AB = AA_full.slice_axis(AA_full.ndim - 2, 0, Dm).slice_axis(AA_full.ndim - 1, Dm, E)
                kn, ka, kz = pk[npl - nk:], pk[npl - ak:], pk[npl - (z.ndim - 2):]
                t1 = D.sum_product([nu, B, z], [kn + ["n", "d"], ka + ["n", "d", "k"], kz + ["n", "k"]], pk + ["n", "d"])
                t2 = D.sum_product([nu, AB, z], [kn + ["n", "d"], ka + ["n", "d", "i", "k"], kz + ["n", "k"]],
                                   pk + ["n", "i"], scale=-1.0)
                nxt, cur = phi0.slice_axis(npl, 1, N), phi0.slice_axis(npl, 0, N - 1)
                D._ew("ADD", nxt.shape, nxt, [nxt, t1])
                D._ew("ADD", cur.shape, cur, [cur, t2])
        return [phi0, phi1, phi2]

    # -- E[log normaliser of the prior] (:251-267, :629-657)
    def compute_cgf_from_parents(self, u_mu, u_Lambda, u_A, u_nu, *u_inputs):
        if self._single is not None and len(u_A) < 4:
            return self._single.compute_cgf_from_parents(u_mu, u_Lambda, u_A, u_nu)
        npl = len(self.plates)
        _, mumu, Lam, logdetL, _, AA_full, nu, lognu = self._parents(u_mu, u_Lambda, u_A, u_nu)
        scale = self._dynamics_scale(u_A)
        if scale is not None:
            lognu = D.add(lognu, scale[1])
        pk = _pk(npl)
        kL, km = pk[npl - (Lam.ndim - 2):], pk[npl - (mumu.ndim - 2):]
        out_pl = pk[npl - max(Lam.ndim - 2, mumu.ndim - 2):]
        t = D.sum_product([Lam, mumu], [kL + ["i", "j"], km + ["i", "j"]], out_pl)
        g = D.axpby(-0.5, t, 0.5, logdetL)
        kn = pk[npl - (lognu.ndim - 2):]
        per_step = 0.5 * ((self.N - 1) if lognu.shape[-2] == 1 else 1.0)      # a time-invariant nu counts N-1 times
        s = D.sum_product([lognu], [kn + ["n", "d"]], kn, scale=per_step)
        g = D.add(g, s)
        if self.K:
            # -1/2 sum_n sum_d nu_d tr(<b_d b_d^T> <z_n z_n^T>)  (:638-655); time-invariant factors count N-1 times
            _, zz = self._inputs(u_inputs)
            Dm, E = self.D, self.E
            BB = AA_full.slice_axis(AA_full.ndim - 2, Dm, E).slice_axis(AA_full.ndim - 1, Dm, E)
            kn, ka, kz = pk[npl - (nu.ndim - 2):], pk[npl - (BB.ndim - 4):], pk[npl - (zz.ndim - 3):]
            out_pl = max((kn, ka, kz), key=len)
            gi = D.sum_product([nu, BB, zz], [kn + ["n", "d"], ka + ["n", "d", "k", "l"], kz + ["n", "k", "l"]],
                               out_pl, scale=-0.5, sizes={"n": self.N - 1})
            g = D.add(g, gi)
        return g

    # -- smoother (:89-123)
    def compute_moments_and_cgf(self, phi, mask=True):
        if self._single is not None:
            return self._single.compute_moments_and_cgf(phi, mask=mask)
        phi = [D.asarray(v) for v in phi]
        N, Dm, P = self.N, self.D, self.plates
        npl = len(P)
        batch = int(np.prod(P, dtype=np.int64)) if P else 1
        be = _bpk.get()
        y = phi[0].broadcast_to(P + (N, Dm)).contiguous()
        A = D.mul(phi[1].broadcast_to(P + (N, Dm, Dm)), -2.0)
        B = D.mul(phi[2].broadcast_to(P + (N - 1, Dm, Dm)), -1.0) if N > 1 else DArray.empty((1,))
        V, x = DArray.empty(P + (N, Dm, Dm)), DArray.empty(P + (N, Dm))
        C = DArray.empty(P + (max(N - 1, 1), Dm, Dm))
        ld = DArray.empty(P)
        be.block_banded_solve(A.ptr, B.ptr, y.ptr, batch, N, Dm, V.ptr, C.ptr, x.ptr, ld.ptr, True)
        u1 = D.add(V, D.mul(x.add_trailing(1), x.expand_dims(-2)))
        if N > 1:
            xp = x.slice_axis(npl, 0, N - 1)
            xn = x.slice_axis(npl, 1, N)
            u2 = D.add(C.slice_axis(npl, 0, N - 1), D.mul(xp.add_trailing(1), xn.expand_dims(-2)))
        else:
            u2 = DArray.empty(P + (0, Dm, Dm))
        pk = _pk(npl)
        g = D.axpby(-0.5, D.sum_product([x, y], [pk + ["n", "i"], pk + ["n", "i"]], pk), 0.5, ld)
        return [x, u1, u2], g

    def compute_fixed_moments_and_f(self, x, mask=True):
        x = np.asarray(x, dtype=np.float64)
        if x.shape[-2:] != (self.N, self.D):
            raise ValueError("Invalid shape")
        u1 = x[..., :, None] * x[..., None, :]
        u2 = x[..., :-1, :, None] * x[..., 1:, None, :]
        return [D.asarray(x), D.asarray(u1), D.asarray(u2)], -0.5 * self.N * self.D * LOG2PI

    # -- messages (:443-527 combined with gaussian.py:2351-2371 / :2496-2522)
    def compute_message_to_parent(self, parent, index, u, u_mu, u_Lambda, u_A, u_nu, *u_inputs):
        gg_dynamics = getattr(parent, "moment_kind", None) == "gaussian_gamma" if index == 2 else \
            (u_A is not None and len(u_A) >= 4)
        if self._single is not None and not gg_dynamics:
            return self._single.compute_message_to_parent(parent, index, u, u_mu, u_Lambda, u_A, u_nu)
        N, Dm, P = self.N, self.D, self.plates
        npl = len(P)
        pk = _pk(npl)
        x, xx, xpxn = u
        mu, mumu, Lam, _, A, AA, nu, _ = self._parents(u_mu, u_Lambda, u_A, u_nu)
        scale = self._dynamics_scale(u_A)
        if index in (0, 1):
            x0 = x.slice_axis(npl, 0, 1).reshape(P + (Dm,))
            x0x0 = xx.slice_axis(npl, 0, 1).reshape(P + (Dm, Dm))
            if index == 0:
                kL = pk[npl - (Lam.ndim - 2):]
                m0 = D.sum_product([Lam, x0], [kL + ["i", "j"], pk + ["j"]], pk + ["i"])
                return [m0, D.mul(Lam, -0.5)]
            xm = D.mul(x0.add_trailing(1), mu.expand_dims(-2))           # x0 mu^T, P + (D, D)
            t = D.add(D.sub(D.sub(x0x0, xm), xm.swap_last2()), mumu)
            return [D.mul(t, -0.5), D.asarray(0.5)]
        if N < 2:
            return [None, None]
        z, zz = self._inputs(u_inputs) if index != 4 else (None, None)
        if index == 4:
            # to the input signals (:504-527): [sum_d nu_d b_d x_{n+1,d} - sum_d nu_d <b_d a_d^T> x_n, -1/2 sum_d nu_d <b_d b_d^T>]
            E = self.E
            B = A.slice_axis(A.ndim - 1, Dm, E)
            AB = AA.slice_axis(AA.ndim - 2, 0, Dm).slice_axis(AA.ndim - 1, Dm, E)
            BB = AA.slice_axis(AA.ndim - 2, Dm, E).slice_axis(AA.ndim - 1, Dm, E)
            kn, ka = pk[npl - (nu.ndim - 2):], pk[npl - (A.ndim - 3):]
            xn, xh = x.slice_axis(npl, 1, N), x.slice_axis(npl, 0, N - 1)
            a = D.sum_product([nu, B, xn], [kn + ["n", "d"], ka + ["n", "d", "k"], pk + ["n", "d"]], pk + ["n", "k"])
            b = D.sum_product([nu, AB, xh], [kn + ["n", "d"], ka + ["n", "d", "i", "k"], pk + ["n", "i"]], pk + ["n", "k"])
            out_pl = max((kn, ka), key=len)
            m1 = D.sum_product([nu, BB], [kn + ["n", "d"], ka + ["n", "d", "k", "l"]], out_pl + ["n", "k", "l"], scale=-0.5)
            return [D.sub(a, b), m1]
        # moments of the regressor c_n = [x_n; z_n] in place of the chain's own (identical without inputs)
        xpxn, xx_head = self._regressor_moments(u, z, zz)
        tau_A = None if scale is None else scale[0]
        if self.static:
            # time sums first (plates kept, time axis of length 1): nothing of size (N-1, D, D, D) is formed
            t1 = ["t"]
            Sxx_head = D.sum_product([xx_head], [pk + ["n", "i", "j"]], pk + t1 + ["i", "j"])
            Sxpxn = D.sum_product([xpxn], [pk + ["n", "i", "j"]], pk + t1 + ["i", "j"])
            if index == 2:
                # to a_d: [nu_d sum_n <x_{n+1,d} x_n>, -1/2 nu_d sum_n <x_n x_n^T>]   plates P + (1, D)
                m0 = D.mul(Sxpxn.swap_last2(), nu.add_trailing(1))
                m1 = D.mul(D.mul(Sxx_head.expand_dims(-3), nu.add_trailing(2)), -0.5)
                if gg_dynamics:
                    # the scale part of the message to Gaussian-gamma dynamics: [-1/2 nu_d sum_n <x_{n+1,d}^2>, (N-1)/2]
                    dxx = D.sum_product([xx.slice_axis(npl, 1, N).diag_view(1)], [pk + ["n", "d"]], pk + t1 + ["d"], scale=-0.5)
                    return [m0, m1, D.mul(dxx, nu), DArray.full(P + (1, Dm), 0.5 * (N - 1))]
                return [m0, m1]
            if index == 3:
                ka = pk[npl - (A.ndim - 3):]
                dxx = D.sum_product([xx.slice_axis(npl, 1, N).diag_view(1)], [pk + ["n", "d"]], pk + t1 + ["d"], scale=-0.5)
                if tau_A is not None:
                    dxx = D.mul(dxx, tau_A)
                a = D.sum_product([Sxpxn, A], [pk + t1 + ["i", "d"], ka + t1 + ["d", "i"]], pk + t1 + ["d"])
                b = D.sum_product([Sxx_head, AA], [pk + t1 + ["i", "j"], ka + t1 + ["d", "i", "j"]], pk + t1 + ["d"], scale=-0.5)
                return [D.add(D.add(dxx, a), b), DArray.full(P + (1, Dm), 0.5 * (N - 1))]
            raise ValueError("Parent index out of bounds")
        # time-varying dynamics: the messages keep the time axis (plates P + (N-1, D))
        if index == 2:
            m0 = D.mul(xpxn.swap_last2(), nu.add_trailing(1))                                   # nu_{n,d} <x_{n+1,d} x_n>
            m1 = D.mul(D.mul(xx_head.expand_dims(-3), nu.add_trailing(2)), -0.5)                # -1/2 nu_{n,d} <x_n x_n^T>
            if gg_dynamics:
                dxx = D.mul(xx.slice_axis(npl, 1, N).diag_view(1), -0.5)
                return [m0, m1, D.mul(dxx, nu), DArray.full(P + (N - 1, Dm), 0.5)]
            return [m0, m1]
        if index == 3:
            ka = pk[npl - (A.ndim - 3):]
            dxx = D.mul(xx.slice_axis(npl, 1, N).diag_view(1), -0.5)
            if tau_A is not None:
                dxx = D.mul(dxx, tau_A)
            a = D.sum_product([xpxn, A], [pk + ["n", "i", "d"], ka + ["n", "d", "i"]], pk + ["n", "d"])
            b = D.sum_product([xx_head, AA], [pk + ["n", "i", "j"], ka + ["n", "d", "i", "j"]], pk + ["n", "d"], scale=-0.5)
            return [D.add(D.add(dxx, a), b), DArray.full(P + (N - 1, Dm), 0.5)]
        raise ValueError("Parent index out of bounds")


class GaussianMarkovChain(ExponentialFamily):
    """``GaussianMarkovChain(mu, Lambda, A, nu, n=None, name="")`` (gaussian_markov_chain.py:660-927)."""
    moment_kind = "gaussian_markov_chain"

    def __init__(self, mu, Lambda, A, nu, n=None, inputs=None, plates=None, name="", initialize=True):
        Lambda = ensure_wishart(Lambda)
        Dm = Lambda.dims[0][-1]
        mu = ensure_gaussian(mu, 1)
        # the dynamics may carry a gamma scale of their own (a Gaussian-gamma node, gaussian_markov_chain.py:817)
        if not (isinstance(A, Node) and A.moment_kind == "gaussian_gamma"):
            A = ensure_gaussian(A, 1)
        nu = ensure_gamma(nu)
        K = 0
        if inputs is not None:
            inputs = ensure_gaussian(inputs, 1)
            if len(inputs.dims[0]) != 1:
                raise ValueError("Input signals have wrong dimensionality")
            K = int(inputs.dims[0][0])
        if tuple(A.dims[0]) != (Dm + K,):
            raise ValueError("Dynamics matrix has wrong dimensionality: rows of length %d expected" % (Dm + K))
        if tuple(A.plates[-1:]) != (Dm,):
            raise ValueError("Dynamics matrix should have a last plate equal to the dimensionality of the system: "
                             "plates (..., N-1 or 1, D), shape (D,)")
        # time extents of the dynamics (gaussian_markov_chain.py:840-880): the second-last plate of A / nu
        TA = int(A.plates[-2]) if len(A.plates) >= 2 else 1
        Tn = int(nu.plates[-2]) if len(nu.plates) >= 2 else 1
        Tz = int(inputs.plates[-1]) if inputs is not None and len(inputs.plates) >= 1 else 1
        n_parents = max(TA, Tn, Tz)
        if len({t for t in (TA, Tn, Tz) if t != 1}) > 1:
            raise ValueError("Plates of parents are giving different number of time instances")
        if n is None:
            if n_parents == 1:
                raise ValueError("The number of time instances could not be determined automatically. "
                                 "Give the number of time instances.")
            n = n_parents + 1
        elif n_parents != 1 and n_parents + 1 != int(n):
            raise ValueError("The number of time instances must match the number of last plates of parents: "
                             "%d != %d+1" % (int(n), n_parents))
        self.N, self.D = int(n), int(Dm)
        from .node import broadcast_plates
        chain_plates = broadcast_plates(tuple(mu.plates), tuple(Lambda.plates), tuple(A.plates[:-2]),
                                        tuple(nu.plates[:-2]), tuple(inputs.plates[:-1]) if inputs is not None else ())
        if plates is not None:
            plates = tuple(int(p) for p in plates)
            chain_plates = broadcast_plates(chain_plates, plates)
            if chain_plates != plates:
                raise ValueError("The plates %s of the parents are not broadcastable to the given plates %s."
                                 % (chain_plates, plates))
        dist = GaussianMarkovChainDistribution(self.N, self.D, chain_plates, TA=TA, Tn=Tn,
                                               plain=len(A.plates) == 1 and len(nu.plates) <= 1 and
                                               A.moment_kind == "gaussian", K=K)
        parents = (mu, Lambda, A, nu) if inputs is None else (mu, Lambda, A, nu, inputs)
        super().__init__(*parents, dims=((self.N, Dm), (self.N, Dm, Dm), (self.N - 1, Dm, Dm)),
                         distribution=dist, plates=chain_plates, name=name, initialize=initialize)

    def _to_gaussian(self):
        if getattr(self, "_as_gaussian", None) is None:
            self._as_gaussian = _MarkovChainToGaussian(self, name=self.name)
        return self._as_gaussian

    def random(self):
        raise NotImplementedError("Sampling from a Gaussian Markov chain is not implemented")

    def rotate(self, R, inv=None, logdet=None):
        """q(x_1..x_N) -> q(R x_1..R x_N)  (gaussian_markov_chain.py:51-65, :167-184): moments by R, natural parameters
        by R^-T, log-normaliser by -N log|det R|.  R is a host D x D matrix; the plated arrays rotate on the device."""
        R = np.asarray(R, dtype=np.float64)
        invR = np.linalg.inv(R) if inv is None else np.asarray(inv, dtype=np.float64)
        logdetR = np.linalg.slogdet(R)[1] if logdet is None else float(logdet)
        Dm = self.D
        Rd, iRT = D.asarray(R), D.asarray(np.ascontiguousarray(invR.T))

        def rot_vec(a, Mx):         # a[..., i] <- sum_k Mx[i, k] a[..., k]
            a = D.asarray(a)
            flat = a.reshape((-1, Dm))
            return D.sum_product([Mx, flat], [["i", "k"], ["n", "k"]], ["n", "i"]).reshape(a.shape)

        def rot_mat(a, Mx):         # a[..., i, j] <- sum_kl Mx[i, k] a[..., k, l] Mx[j, l]
            a = D.asarray(a)
            flat = a.reshape((-1, Dm, Dm))
            t = D.sum_product([Mx, flat], [["i", "k"], ["n", "k", "l"]], ["n", "i", "l"])
            return D.sum_product([t, Mx], [["n", "i", "l"], ["j", "l"]], ["n", "i", "j"]).reshape(a.shape)

        self.u = [rot_vec(self.u[0], Rd), rot_mat(self.u[1], Rd), rot_mat(self.u[2], Rd)]
        self.phi = [rot_vec(self.phi[0], iRT), rot_mat(self.phi[1], iRT), rot_mat(self.phi[2], iRT)]
        self.g = D.affine(D.asarray(self.g), 1.0, -self.N * logdetR)
        self._version += 1


class _AddTrailingPlate(Deterministic):
    """The parent seen with one more (unit) plate axis: after its own plates (``position`` 0) or in front of its last
    ``position`` plates.  Moments and messages are the parent's, reshaped.  Lets a time-plated node broadcast against a
    node whose last plate is the state dimension, and plated mixing matrices (..., D) against the time axis."""

    def __init__(self, node, name="", position=0):
        self.moment_kind = node.moment_kind
        self.position = int(position)
        if self.position > len(node.plates):
            raise ValueError("The node has fewer than %d plates" % self.position)
        super().__init__(node, dims=node.dims, plates=self._insert(tuple(node.plates)), name=name)

    def _insert(self, plates, value=1):
        k = len(plates) - self.position
        return tuple(plates[:k]) + (value,) + tuple(plates[k:])

    def _plates_from_parent(self, index):
        return self._insert(tuple(self.parents[0].plates))

    def _map_parent_axes(self, index, values):
        return self._insert(tuple(values))

    def _plates_to_parent(self, index):
        k = len(self.plates) - 1 - self.position
        return tuple(self.plates[:k]) + tuple(self.plates[k + 1:])

    def _weights_to_parent(self, index, mask):
        mask = np.asarray(mask)
        return np.any(mask, axis=-1 - self.position) if mask.ndim >= 1 + self.position else mask

    def _compute_moments(self, u):
        out = []
        for ui, dims in zip(u, self.dims):
            ui = D.asarray(dense(ui))
            npl = ui.ndim - len(dims)
            if npl >= self.position:
                k = npl - self.position
                ui = ui.reshape(tuple(ui.shape[:k]) + (1,) + tuple(ui.shape[k:]))
            out.append(ui)
        return out

    def message_to_parent(self, index):
        # the children have summed their messages to this node's plates: drop the unit axis
        m = self.message_from_children()
        out = []
        for mi, dims in zip(m, self.dims):
            if mi is None:
                out.append(None)
                continue
            mi = D.asarray(mi)
            npl = mi.ndim - len(dims)
            if npl >= 1 + self.position:
                k = npl - 1 - self.position
                mi = mi.reshape(tuple(mi.shape[:k]) + tuple(mi.shape[k + 1:]))
            out.append(mi)
        return out


class VaryingGaussianMarkovChain(GaussianMarkovChain):
    """``VaryingGaussianMarkovChain(mu, Lambda, B, S, nu, n=None, name="")`` (gaussian_markov_chain.py:1200-1452): a
    chain whose transition matrix mixes K matrices with time-varying weights, A_n = sum_k s_nk B_k, with B of shape
    (D, K) and plates (..., D) and the weights S of shape (K,) and plates (..., N-1).

    The reference has a dedicated distribution for it and notes that the same model is a ``GaussianMarkovChain`` over
    ``SumMultiply`` (``:1296-1301``); that is how it is built here: the per-step dynamics moments <A_n>, <a_nd a_nd^T>
    are one device contraction each, and the messages to B and S are the chain's messages to its dynamics carried
    through the product node."""

    def __init__(self, mu, Lambda, B, S, nu, n=None, plates=None, name="", initialize=True):
        from .dot import SumMultiply
        from .gaussian import ensure_gaussian as _eg
        B = _eg(B, 2)
        S = _eg(S, 1)
        if len(B.dims[0]) != 2:
            raise ValueError("Third parent has wrong dimensionality {0}.".format(B.dims[0]))
        Dm, K = B.dims[0]
        if len(B.plates) == 0 or B.plates[-1] != Dm:
            raise ValueError("Third parent should have a last plate equal to the dimensionality of the system.")
        if tuple(S.dims[0]) != (K,):
            raise ValueError("Fourth parent has wrong dimensionality %s, should be %s" % (S.dims, ((K,), (K, K))))
        nS = S.plates[-1] if len(S.plates) >= 1 else 1
        if n is not None and nS != 1 and nS != int(n) - 1:
            raise ValueError("The last plate of the fourth parent should have length equal to one or N-1, where N is "
                             "the number of time instances.")
        self._mixing = (B, S)
        weights = _AddTrailingPlate(S) if len(S.plates) >= 1 else S
        # chains over plates: the mixing matrices (..., D) meet the time axis of the weights (..., N-1, 1) as (..., 1, D)
        mixing = _AddTrailingPlate(B, position=1) if len(B.plates) >= 2 else B
        A = SumMultiply("jk,k->j", mixing, weights)
        super().__init__(mu, Lambda, A, nu, n=n, plates=plates, name=name, initialize=initialize)


class SwitchingGaussianMarkovChain(GaussianMarkovChain):
    """``SwitchingGaussianMarkovChain(mu, Lambda, B, Z, nu, n=None, name="")`` (gaussian_markov_chain.py:1743-1985): at
    every step one of K transition matrices is selected, A_n = B_{z_n}; B has plates (..., K, D) and D-dimensional rows,
    Z is categorical with plates (..., N-1).  As the reference notes (``:1846-1851``) this is a ``GaussianMarkovChain``
    over ``Gate``, which is how it is built here."""

    def __init__(self, mu, Lambda, B, Z, nu, n=None, plates=None, name="", initialize=True):
        from .gate import Gate
        from .categorical import categorical_constant
        B = ensure_gaussian(B, 1)
        if len(B.plates) < 2:
            raise ValueError("Third parent should have plates (..., K, D)")
        K, Dm = int(B.plates[-2]), int(B.plates[-1])
        if tuple(B.dims[0]) != (Dm,):
            raise ValueError("Third parent should have a last plate equal to the dimensionality of the system.")
        if isinstance(Z, Node) and hasattr(Z, "_to_categorical"):
            Z = Z._to_categorical()
        if not isinstance(Z, Node):
            Z = categorical_constant(Z, K)
        if Z.moment_kind != "categorical" or tuple(Z.dims) != ((K,),):
            raise ValueError("Fourth parent has wrong dimensionality: %d categories expected" % K)
        nZ = Z.plates[-1] if len(Z.plates) >= 1 else 1
        if n is not None and nZ != 1 and nZ != int(n) - 1:
            raise ValueError("The last plate of the fourth parent should have length equal to one or N-1, where N is "
                             "the number of time instances.")
        self._switching = (B, Z)
        selector = _AddTrailingPlate(Z) if len(Z.plates) >= 1 else Z
        A = Gate(selector, B, gated_plate=-2)
        super().__init__(mu, Lambda, A, nu, n=n, plates=plates, name=name, initialize=initialize)


class _MarkovChainToGaussian(Deterministic):
    """The chain seen as N Gaussian vectors plated over time (gaussian_markov_chain.py:1988-2098): the
    time axis of the parent's dims is the last plate here; the cross-time moment is not exposed."""
    moment_kind = "gaussian"

    def __init__(self, X, name=""):
        self.N, self.D = X.N, X.D
        super().__init__(X, dims=((X.D,), (X.D, X.D)), plates=tuple(X.plates) + (X.N,), name=name)

    def _plates_from_parent(self, index):
        return tuple(self.parents[0].plates) + (self.N,)

    def _plates_to_parent(self, index):
        return tuple(self.plates[:-1])

    def _weights_to_parent(self, index, mask):
        mask = np.asarray(mask)
        return np.any(mask, axis=-1) if mask.ndim >= 1 else mask

    def _compute_moments(self, u):
        return [u[0], u[1]]

    def message_to_parent(self, index):
        # every child masks its own message, so the time axis can simply turn from plate into dim
        m = self.message_from_children()
        out = []
        for i, mi in enumerate(m):
            if mi is None:
                out.append(None)
                continue
            want = (self.N,) + (self.D,) * (i + 1)
            mi = D.asarray(mi)
            if mi.ndim < len(want):
                mi = mi.add_leading(len(want) - mi.ndim)
            out.append(mi)
        return out + [None]
