"""Fused sweep plans: sub-graphs that one pass over the data can serve.

The reference recomputes every message on every request (node.py:657-688,
deterministic.py:72-82), so a PCA sweep traverses the (M, N) data >= 6 times
and materialises (N, K, K) second moments.  A *plan* recognises a sub-graph,
and re-routes the handful of node methods whose cost scales with the plate
count to fused kernels and to cached plate-summed sufficient statistics:

FactorModelPlan   Y ~ GaussianARD(SumMultiply('k,k->', A, B), tau), observed,
                  A plated (M,1), B plated (1,N)  (Bayesian PCA, pca.rst:40-66)

    B.update()                     -> bpk_pca_xsweep: one pass over Y writes <b_n>
                                      and accumulates S_yx, S_xx, s_x
    F.message_to_parent(A)         -> from the statistics        (dot.py:581 sums over n)
    Y.message_to_parent(tau)       -> from the statistics        (gaussian.py:2361-2369)
    Y.lower_bound_contribution()   -> from the statistics        (expfamily.py:400-480)
    B.lower_bound_contribution()   -> from the statistics

The statistics are tagged with the version counter of the node they were
computed from, so any update order the user chooses stays correct: a stale tag
triggers ``bpk_pca_stats`` (one pass over Y and <B>) instead of reusing them.
Whenever a precondition does not hold (missing values, annealing, plated tau,
extra children ...) the plan steps aside and the generic path runs.
"""
import types

import numpy as np

from .. import _bpk
from .. import darray as D
from .. import parallel
from ..darray import DArray
from .dot import SumMultiply
from .expfam import ExponentialFamily
from .gamma import Gamma
from .gaussian import GaussianARD, FactoredSecondMoment, dense, LOG2PI
from .node import Constant, mask_is_full


class LazyArray:
    """Device array produced on first use (kept out of the sweep's HBM traffic)."""

    def __init__(self, shape, fn):
        self.shape = tuple(shape)
        self._fn = fn
        self._val = None

    def materialize(self):
        if self._val is None:
            self._val = self._fn()
            self._fn = None
        return self._val

    @property
    def ndim(self):
        return len(self.shape)

    def numpy(self):
        return self.materialize().numpy()

    def __array__(self, dtype=None, copy=None):
        return self.numpy()


def _scalar_plates(node):
    return all(n == 1 for n in node.plates)


class FactorModelPlan:

    def __init__(self, Y, F, row, col, i_row, i_col, tau):
        self.Y, self.F, self.row, self.col, self.tau = Y, F, row, col, tau
        self.i_row, self.i_col = i_row, i_col
        self.M, self.N = Y.plates
        self.K = row.dims[0][0]
        # plate sharding: N is this rank's block; Ng the global column count
        self.world = parallel.world()
        self.Ng = int(round(float(np.sum(parallel.allgather_scalar(self.N))))) if self.world > 1 else self.N
        self.kernel_timers = None     # bench.py: list of timer ids to record around the sweep kernel
        self.kernel_timer_log = []    # [(timer ids used by a chunk, sweeps the chunk ran)]
        self._timer_pos = 0
        self._stats = None            # (col._version, DArray)
        self._sumsq = None            # (Y._version, DArray[2])
        self._e2 = None               # ((row._version, col._version), DArray scalar)
        self._orig = {}
        self.fused_calls = 0
        self._install()

    # ---- wiring ---------------------------------------------------------------------------------
    def _install(self):
        plan = self
        o = self._orig
        o["col.update"] = self.col.update
        o["Y.update"] = self.Y.update
        o["col.lb"] = self.col.lower_bound_contribution
        o["F.msg"] = self.F.message_to_parent
        o["Y.msg"] = self.Y.message_to_parent
        o["Y.lb"] = self.Y.lower_bound_contribution

        def col_update(node, annealing=1.0):
            if annealing == 1.0 and plan.valid():
                if plan.masked():
                    return plan.update_col_masked()
                return plan.update_col()
            return o["col.update"](annealing) if annealing != 1.0 else o["col.update"]()

        def Y_update(node, annealing=1.0):
            if annealing == 1.0 and plan.valid() and plan.masked():
                return plan.update_Y_lazy()
            return o["Y.update"](annealing) if annealing != 1.0 else o["Y.update"]()

        def col_lb(node, ignore_masked=True):
            if not ignore_masked:
                plan._materialize_Y()
                return o["col.lb"](ignore_masked=False)
            if plan.valid():
                if plan.masked():
                    if plan._mstats_ready():
                        return plan.bound_col_masked()
                elif plan._col_is_fused():
                    return plan.bound_col()
            plan._materialize_Y()
            return o["col.lb"]()

        def F_msg(node, index):
            if index == plan.i_row and plan.valid():
                if not plan.masked():
                    return plan.message_to_row()
                if plan._mstats_ready():
                    return plan.message_to_row_masked()
            plan._materialize_Y()
            return o["F.msg"](index)

        def Y_msg(node, index):
            if index == 1 and plan.valid():
                if not plan.masked():
                    return plan.message_to_tau()
                if plan._mstats_ready():
                    return plan.message_to_tau_masked()
            plan._materialize_Y()
            return o["Y.msg"](index)

        def Y_lb(node, ignore_masked=True):
            if not ignore_masked:
                plan._materialize_Y()
                return o["Y.lb"](ignore_masked=False)
            if plan.valid():
                if not plan.masked():
                    return plan.bound_Y()
                if plan._mstats_ready():
                    return plan.bound_Y_masked()
            plan._materialize_Y()
            return o["Y.lb"]()

        def col_rotated(Rd, old_version):
            # S_yx = sum y x^T, S_xx = sum x x^T, s_x = sum x follow x -> R x without another pass over Y
            if plan._stats is not None and plan._stats[0] == old_version:
                M, K = plan.M, plan.K
                Syx, Sxx, sx = plan._split(plan._stats[1])
                st = DArray.empty((M * K + K * K + K,))
                a, b, c = plan._split(st)
                D.sum_product([Syx, Rd], [["m", "k"], ["i", "k"]], ["m", "i"], out=a)
                D.sum_product([Rd, Sxx, Rd], [["i", "k"], ["k", "l"], ["j", "l"]], ["i", "j"], out=b)
                D.sum_product([Rd, sx], [["i", "k"], ["k"]], ["i"], out=c)
                plan._stats = (plan.col._version, st)
            plan._e2 = None
            plan._res_cache = None
        hooks = list(getattr(self.col, "_rotate_hooks", ()))
        hooks.append(col_rotated)
        self.col._rotate_hooks = hooks

        self.col.update = types.MethodType(col_update, self.col)
        self.Y.update = types.MethodType(Y_update, self.Y)
        self.col.lower_bound_contribution = types.MethodType(col_lb, self.col)
        self.F.message_to_parent = types.MethodType(F_msg, self.F)
        self.Y.message_to_parent = types.MethodType(Y_msg, self.Y)
        self.Y.lower_bound_contribution = types.MethodType(Y_lb, self.Y)

    def valid(self):
        ok = self._valid()
        if not ok and self.world > 1:
            raise NotImplementedError("plate sharding is only supported on the fused path "
                                      "(fully observed data, no annealing)")
        return ok

    def _structure_ok(self):
        """The graph still has the shape ``attach`` matched: nobody hung another child on the factors, the
        product node or the data since (their messages would be dropped by the fused updates).  Cheap; checked on
        every call, because nodes can be added to a model after VB(...) was built."""
        return (len(self.col.children) == 1 and len(self.row.children) == 1 and len(self.F.children) == 1
                and not self.Y.children)

    def _valid(self):
        Y, col, row = self.Y, self.col, self.row
        if not self._structure_ok():
            self._res_cache = None
            return False
        if not (Y.observed is True or (isinstance(Y.observed, np.ndarray) and Y.observed.all())):
            if not self.masked():
                return False
        elif not mask_is_full(Y.mask):
            return False
        if any(n.annealing != 1.0 for n in (Y, col, row) if hasattr(n, "annealing")):
            return False
        if col.observed is not False or row.observed is not False:
            return False
        if any(getattr(n, "plates_multiplier", ()) for n in (Y, self.F, col, row)):
            return False                                   # mini-batch scaling: generic path
        return True

    # ---- missing values: per-column precision, fused (csrc/pca_masked.cu) -----------------------------------------
    def masked(self):
        """True when Y is observed through an (M, N) mask the fused masked sweep can serve: every column and every row
        has at least one observation (so that no plate of X or C is switched off: node.py:446-526), M <= 64, K <= 16."""
        Y = self.Y
        key = (id(Y.observed), id(Y.mask))
        if getattr(self, "_masked_key", None) != key:
            ok = False
            obs = Y.observed
            if isinstance(obs, np.ndarray) and not obs.all() and obs.shape == (self.M, self.N) \
                    and self.M <= 64 and self.K <= 16 and hasattr(_bpk.get(), "pca_xsweep_masked_fused"):
                mk = np.broadcast_to(np.asarray(Y.mask, dtype=bool), (self.M, self.N))
                ok = bool(np.array_equal(mk, obs) and obs.any(axis=0).all() and obs.any(axis=1).all())
            self._masked_ok = ok
            self._masked_key = key
            self._mstats = None
        return self._masked_ok

    def _mstats_ready(self):
        """The masked statistics describe the CURRENT q(X) (they were produced by the sweep that formed it)."""
        ms = getattr(self, "_mstats", None)
        return ms is not None and ms[0] == self.col._version

    def _msplit(self):
        M, K = self.M, self.K
        st = self._mstats[1]
        o = [0, M * K, M * K + M * K * K, M * K + M * K * K + K * K, M * K + M * K * K + K * K + K]
        return (st.slice_axis(0, o[0], o[1]).reshape((M, K)), st.slice_axis(0, o[1], o[2]).reshape((M, K, K)),
                st.slice_axis(0, o[2], o[3]).reshape((K, K)), st.slice_axis(0, o[3], o[4]),
                st.slice_axis(0, o[4], o[4] + 1).reshape(()), st.slice_axis(0, o[4] + 1, o[4] + 2).reshape(()))

    def _row_moments(self):
        M, K = self.M, self.K
        W = self.row.u[0].broadcast_to((M, 1, K)).reshape((M, K)).contiguous()
        WW = dense(self.row.u[1])
        WW = D.asarray(WW).broadcast_to((M, 1, K, K)).reshape((M, K, K)).contiguous()
        return W, WW

    def _sumsq_masked(self):
        v = id(self.Y.observed)
        if getattr(self, "_msumsq", None) is None or self._msumsq[0] != v:
            out = DArray.empty((2,))
            _bpk.get().sumsq(self._Yd().ptr, self.Y.mask_device().contiguous().ptr, self.M * self.N, out.ptr)
            parallel.allreduce_sum(out)
            self._msumsq = (v, out, float(out.numpy()[1]))
        return self._msumsq[1].slice_axis(0, 0, 1).reshape(()), self._msumsq[2]

    def _materialize_Y(self):
        """Generic code is about to read Y's moments: make the pending latent entries concrete."""
        Y = self.Y
        for name in ("u", "phi"):
            lst = getattr(Y, name, None)
            if isinstance(lst, list) and any(isinstance(a, LazyArray) for a in lst):
                setattr(Y, name, [dense(a) for a in lst])
        if isinstance(getattr(Y, "g", None), LazyArray):
            Y.g = dense(Y.g)

    def update_Y_lazy(self):
        """Y.update() for a partially observed Y (stochastic.py:276-282: the unobserved entries of Y are latent and get
        the predictive moments of their parents).  Nothing downstream of the sweep uses them — messages and the bound
        only see the observed entries — so they are produced ON DEMAND (``np.asarray(Y.u[0])``), from the parents'
        moments as they are NOW (the arrays are captured; forming them needs <f f> over (M, N) and the per-column
        second moments, i.e. exactly the passes the fused sweep avoids)."""
        Y, F = self.Y, self.F
        snaps = [list(p.u) for p in F.parents]
        utau = list(self.tau.get_moments())
        ydata, obs = self._Yd(), Y.mask_device()
        memo = {}

        def compute():
            if not memo:
                uF = F._compute_moments(*[[D.asarray(u[0]), dense(u[1])] for u in snaps])
                phi = Y._canonical_phi(Y._distribution.compute_phi_from_parents(uF, utau))
                u_new, g_new = Y._distribution.compute_moments_and_cgf(phi)
                fixed = [ydata, D.square(ydata)]
                memo["u"] = [D.where(obs.add_trailing(Y.ndims[i]), fixed[i], D.asarray(u_new[i]).broadcast_to(Y.get_shape(i)))
                             for i in range(2)]
                memo["phi"] = [D.asarray(p) for p in phi]
                memo["g"] = D.asarray(g_new)
            return memo
        Y.u = [LazyArray(Y.get_shape(0), lambda: compute()["u"][0]), LazyArray(Y.get_shape(1), lambda: compute()["u"][1])]
        Y.phi = [LazyArray(Y.get_shape(0), lambda: compute()["phi"][0]), LazyArray(Y.get_shape(1), lambda: compute()["phi"][1])]
        Y.g = LazyArray(tuple(Y.plates), lambda: compute()["g"])
        Y._version += 1

    def update_col_masked(self):
        """col.update() with missing values: one fused sweep builds, inverts and contracts the per-column precisions
        (gaussian.py:672-706 + linalg.py:50-59,111-146,185-195 + dot.py:581 in the reference)."""
        col, M, N, K = self.col, self.M, self.N, self.K
        be = _bpk.get()
        u_par = col.moments_from_parents()
        phi_p = col._canonical_phi(col._distribution.compute_phi_from_parents(*u_par))
        if not (all(n == 1 for n in phi_p[0].shape[:-1]) and all(n == 1 for n in phi_p[1].shape[:-2])):
            return self._orig["col.update"]()          # prior differs per column: generic path
        tau = float(self._tau_moments()[0].numpy())
        W, WW = self._row_moments()
        amu = phi_p[0].reshape((K,)).contiguous()
        alpha = D.mul(phi_p[1].reshape((K, K)).diag_view(1), -2.0).contiguous()        # prior precision (diagonal)
        X = DArray.empty((1, N, K))
        g = DArray.empty((1, N))
        st = DArray.zeros((M * K + M * K * K + K * K + K + 2,))
        Yd, mk = self._Yd(), self.Y.mask_device().contiguous()
        if self.kernel_timers is not None and self._timer_pos < len(self.kernel_timers):
            tid = self.kernel_timers[self._timer_pos]
            self._timer_pos += 1
            be.timer_record(tid, 0)
            be.pca_xsweep_masked_fused(Yd.ptr, mk.ptr, M, N, K, W.ptr, WW.ptr, tau, alpha.ptr, amu.ptr, X.ptr, g.ptr, st.ptr)
            be.timer_record(tid, 1)
        else:
            be.pca_xsweep_masked_fused(Yd.ptr, mk.ptr, M, N, K, W.ptr, WW.ptr, tau, alpha.ptr, amu.ptr, X.ptr, g.ptr, st.ptr)
        parallel.allreduce_sum(st)
        self.fused_calls += 1

        def cov_fn():
            """(1,N,K,K) covariances on demand only (the sweep itself never stores them)."""
            cov = DArray.empty((1, N, K, K))
            Xs, gs = DArray.empty((N, K)), DArray.empty((N,))
            s2 = DArray.zeros((M * K + M * K * K,))
            be.pca_xsweep_masked(Yd.ptr, mk.ptr, M, N, K, W.ptr, WW.ptr, tau, alpha.ptr, amu.ptr, Xs.ptr, cov.ptr, gs.ptr,
                                 s2.ptr)
            return cov
        lazy_cov = LazyArray((1, N, K, K), cov_fn)

        def phi1_fn():
            c = dense(lazy_cov)
            out = DArray.empty((1, N, K, K))
            # -1/2 Lam_n = -1/2 Cov_n^-1: re-inverting is only for users who look at phi; the bound does not need it
            U = DArray.empty((N, K, K))
            be.chol(c.reshape((N, K, K)).contiguous().ptr, U.ptr, N, K)
            be.chol_inv(U.ptr, out.ptr, N, K)
            return D.mul(out, -0.5)

        def phi0_fn():
            Lam = D.mul(phi1_fn(), -2.0)
            return D.sum_product([Lam, X], [["o", "n", "i", "j"], ["o", "n", "j"]], ["o", "n", "i"])
        col.phi = [LazyArray((1, N, K), phi0_fn), LazyArray((1, N, K, K), phi1_fn)]
        col.u = [X, FactoredSecondMoment(X, lazy_cov, (K,))]
        col.g = g
        col._version += 1
        col._fused = None
        self._mstats = (col._version, st)
        self._stats = None

    def message_to_row_masked(self):
        """F -> row factor with missing values: m0[m] = tau S_yx[m], m1[m] = -1/2 tau sum_n mask[m,n] <x_n x_n^T>
        (dot.py:581 masked and summed over n by node.py:650)."""
        tau, _ = self._tau_moments()
        Syx, Sxx, _, _, _, _ = self._msplit()
        M, K = self.M, self.K
        self.fused_calls += 1
        return [D.mul(Syx, tau).reshape((M, 1, K)), D.mul(D.mul(Sxx, tau), -0.5).reshape((M, 1, K, K))]

    def _E2_masked(self):
        """sum over the observed entries of <(y - f)^2>."""
        key = (self.row._version, self.col._version)
        if getattr(self, "_me2", None) is None or self._me2[0] != key:
            Syx, Sxx, _, _, _, _ = self._msplit()
            W, WW = self._row_moments()
            t1 = D.sum_product([W, Syx], [["m", "k"], ["m", "k"]], [])
            t2 = D.sum_product([WW, Sxx], [["m", "i", "j"], ["m", "i", "j"]], [])
            ssq, _ = self._sumsq_masked()
            self._me2 = (key, D.add(D.axpby(1.0, ssq, -2.0, t1), t2))
        return self._me2[1]

    def message_to_tau_masked(self):
        self.fused_calls += 1
        shp = tuple(self.tau.plates)
        _, count = self._sumsq_masked()
        return [D.mul(self._E2_masked(), -0.5).reshape(shp), D.asarray(0.5 * count).reshape(shp)]

    def bound_Y_masked(self):
        """Observed entries only (expfamily.py:470-476 sums where the node's mask is set)."""
        tau, logtau = self._tau_moments()
        _, count = self._sumsq_masked()
        a = D.mul(D.mul(self._E2_masked(), tau), -0.5)
        return D.add(a, D.affine(logtau, 0.5 * count, -0.5 * count * LOG2PI))

    def bound_col_masked(self):
        """E[log p(X)] - E[log q(X)]: the q-side terms reduce to N K / 2 - 1/2 sum_n log det Lam_n; the prior is
        evaluated with the parents' CURRENT moments (expfamily.py:400-480)."""
        col, N, K = self.col, self.Ng, self.K
        u_par = col.moments_from_parents()
        phi_p = col._canonical_phi(col._distribution.compute_phi_from_parents(*u_par))
        if not (all(n == 1 for n in phi_p[0].shape[:-1]) and all(n == 1 for n in phi_p[1].shape[:-2])):
            return self._orig["col.lb"]()
        cgf = D.asarray(col._distribution.compute_cgf_from_parents(*u_par))
        _, _, sumXX, sumx, _, sumld = self._msplit()
        t0 = D.sum_product([phi_p[0].reshape((K,)), sumx], [["k"], ["k"]], [])
        t1 = D.sum_product([phi_p[1].reshape((K, K)), sumXX], [["i", "j"], ["i", "j"]], [])
        ncgf = D.mul(D.reduce_to_shape(cgf, (), from_shape=col.plates), float(self.Ng) / float(self.N))
        return D.add(D.add(D.add(t0, t1), D.affine(sumld, -0.5, 0.5 * N * K)), ncgf)

    def _col_is_fused(self):
        return getattr(self.col, "_fused", None) is not None and isinstance(self.col.u[1], FactoredSecondMoment) and \
            all(n == 1 for n in self.col.u[1].cov.shape[:-2])

    # ---- small shared quantities ---------------------------------------------------------------------
    def _Yd(self):
        u0 = self.Y.u[0]
        if isinstance(u0, LazyArray):            # latent entries pending (update_Y_lazy): the observed values are unchanged
            return self._ydata
        self._ydata = u0.contiguous()
        return self._ydata

    def _tau_moments(self):
        u = self.tau.get_moments()
        return u[0].reshape(()), u[1].reshape(())

    def _sum_second_moment(self, node, count):
        """sum over the node's plates of <x x^T>  -> (K, K)."""
        K = self.K
        u1 = node.u[1]
        if isinstance(u1, FactoredSecondMoment):
            W = u1.u0.reshape((-1, K))
            S = D.sum_product([W, W], [["m", "i"], ["m", "j"]], ["i", "j"])
            cov = u1.cov.reshape((-1, K, K))
            D.sum_product([cov], [["c", "i", "j"]], ["i", "j"], out=S, accumulate=True,
                          scale=float(count) / cov.shape[0])
            return S
        u1 = D.asarray(u1).reshape((-1, K, K))
        return D.sum_product([u1], [["m", "i", "j"]], ["i", "j"], scale=float(count) / u1.shape[0])

    def _mean(self, node):
        return node.u[0].reshape((-1, self.K))

    def stats(self):
        """[S_yx | S_xx | s_x] for the current posterior of the column factor."""
        v = self.col._version
        if self._stats is None or self._stats[0] != v:
            M, N, K = self.M, self.N, self.K
            st = DArray.zeros((M * K + K * K + K,))
            X = self._mean(self.col)
            if X.shape[0] != N:
                X = self.col.u[0].broadcast_to((1, N, K)).reshape((N, K))
            X = X.contiguous()
            _bpk.get().pca_stats(self._Yd().ptr, M, N, K, X.ptr, st.ptr)
            parallel.allreduce_sum(st)
            self._stats = (v, st)
        return self._split(self._stats[1])

    def _split(self, st):
        M, K = self.M, self.K
        return (st.slice_axis(0, 0, M * K).reshape((M, K)),
                st.slice_axis(0, M * K, M * K + K * K).reshape((K, K)),
                st.slice_axis(0, M * K + K * K, M * K + K * K + K))

    def _sum_xx(self):
        """sum_n <x_n x_n^T> = sum_n Cov_n + S_xx."""
        _, Sxx, _ = self.stats()
        u1 = self.col.u[1]
        if isinstance(u1, FactoredSecondMoment):
            cov = u1.cov.reshape((-1, self.K, self.K))
            if cov.shape[0] == 1:
                return D.sum_product([cov], [["c", "i", "j"]], ["i", "j"], out=Sxx.copy(), accumulate=True,
                                     scale=float(self.Ng))
            if self.world > 1:
                raise NotImplementedError("per-column covariances are not supported with plate sharding yet")
            return D.sum_product([cov], [["c", "i", "j"]], ["i", "j"], out=Sxx.copy(), accumulate=True)
        if self.world > 1:
            raise NotImplementedError("dense second moments are not supported with plate sharding yet")
        return self._sum_second_moment(self.col, self.N)

    def _sumsq_y(self):
        v = self.Y._version
        if self._sumsq is None or self._sumsq[0] != v:
            out = DArray.empty((2,))
            _bpk.get().sumsq(self._Yd().ptr, 0, self.M * self.N, out.ptr)
            parallel.allreduce_sum(out)
            self._sumsq = (v, out)
        return self._sumsq[1].slice_axis(0, 0, 1).reshape(())

    def _E2(self):
        """sum_mn <(y - f)^2> = sum y^2 - 2 <W>:S_yx + tr(sum_m<ww^T> sum_n<xx^T>)."""
        key = (self.row._version, self.col._version)
        if self._e2 is None or self._e2[0] != key:
            Syx, _, _ = self.stats()
            W = self._mean(self.row)
            if W.shape[0] != self.M:
                W = self.row.u[0].broadcast_to((self.M, 1, self.K)).reshape((self.M, self.K))
            t1 = D.sum_product([W, Syx], [["m", "k"], ["m", "k"]], [])
            t2 = D.sum_product([self._sum_second_moment(self.row, self.M), self._sum_xx()],
                               [["i", "j"], ["i", "j"]], [])
            e2 = D.add(D.axpby(1.0, self._sumsq_y(), -2.0, t1), t2)
            self._e2 = (key, e2)
        return self._e2[1]

    # ---- fused node operations -----------------------------------------------------------------------
    def update_col(self):
        """col.update(): prior + messages + moments of the (1,N)-plated factor in one pass over Y."""
        col, M, N, K = self.col, self.M, self.N, self.K
        be = _bpk.get()
        u_par = col.moments_from_parents()
        phi_p = col._canonical_phi(col._distribution.compute_phi_from_parents(*u_par))
        if not (all(n == 1 for n in phi_p[0].shape[:-1]) and all(n == 1 for n in phi_p[1].shape[:-2])):
            return self._orig["col.update"]()          # prior differs per column: generic path
        tau, _ = self._tau_moments()
        SWW = self._sum_second_moment(self.row, M)
        phi1 = D.fma(-0.5, tau, SWW, 1.0, phi_p[1])                        # (1,1,K,K)
        phi0p = phi_p[0].reshape((1, K)).contiguous()
        phi1c = phi1.reshape((1, K, K)).contiguous()
        b = DArray.empty((K,))
        cov = DArray.empty((1, 1, K, K))
        logdet = DArray.empty(())
        be.gaussian_moments(phi0p.ptr, 1, phi1c.ptr, 1, 1, K, b.ptr, cov.ptr, 0, logdet.ptr, True)
        W = self._mean(self.row)
        if W.shape[0] != M:
            W = self.row.u[0].broadcast_to((M, 1, K)).reshape((M, K))
        A = D.sum_product([tau, cov.reshape((K, K)), W], [[], ["k", "j"], ["m", "j"]], ["k", "m"])
        X = DArray.empty((1, N, K))
        st = DArray.zeros((M * K + K * K + K,))
        Yd = self._Yd()
        if self.kernel_timers is not None and self._timer_pos < len(self.kernel_timers):
            tid = self.kernel_timers[self._timer_pos]
            self._timer_pos += 1
            be.timer_record(tid, 0)
            be.pca_xsweep(Yd.ptr, M, N, K, A.ptr, b.ptr, X.ptr, st.ptr)
            be.timer_record(tid, 1)
        else:
            be.pca_xsweep(Yd.ptr, M, N, K, A.ptr, b.ptr, X.ptr, st.ptr)
        parallel.allreduce_sum(st)            # the one exchange step of the sweep
        self.fused_calls += 1
        # publish the new posterior; phi0 / g / <xx^T> stay virtual until somebody looks
        Lam = D.mul(phi1.reshape((K, K)), -2.0)

        def phi0_fn():
            return D.sum_product([Lam, X], [["i", "j"], ["o", "n", "j"]], ["o", "n", "i"])

        def g_fn():
            q = D.sum_product([X, Lam, X], [["o", "n", "i"], ["i", "j"], ["o", "n", "j"]], ["o", "n"])
            return D.axpby(-0.5, q, 0.5, logdet)
        col.phi = [LazyArray((1, N, K), phi0_fn), phi1]
        col.u = [X, FactoredSecondMoment(X, cov, (K,))]
        col.g = LazyArray((1, N), g_fn)
        col._version += 1
        col._fused = dict(Lam=Lam, logdet=logdet, phi_p=phi_p, cov=cov)
        self._stats = (col._version, st)

    def message_to_row(self):
        """F -> row factor:  m0[m] = tau S_yx[m],  m1 = -1/2 tau sum_n <x x^T>  (dot.py:581 summed over n)."""
        tau, _ = self._tau_moments()
        Syx, _, _ = self.stats()
        M, K = self.M, self.K
        m0 = D.mul(Syx, tau).reshape((M, 1, K))
        m1 = D.mul(D.mul(self._sum_xx(), tau), -0.5).reshape((1, 1, K, K))
        self.fused_calls += 1
        return [m0, m1]

    def message_to_tau(self):
        """Y -> tau: [-1/2 sum <(y-f)^2>, 1/2 #obs] reduced to tau's plates (gaussian.py:2361-2369)."""
        self.fused_calls += 1
        shp = tuple(self.tau.plates)
        m0 = D.mul(self._E2(), -0.5).reshape(shp)
        m1 = D.asarray(0.5 * self.M * self.Ng).reshape(shp)
        return [m0, m1]

    def bound_Y(self):
        """sum_mn <log N(y | f, 1/tau)> = -1/2 tau E2 + MN/2 (<log tau> - log 2pi)."""
        tau, logtau = self._tau_moments()
        MN = float(self.M) * float(self.Ng)
        a = D.mul(D.mul(self._E2(), tau), -0.5)
        return D.add(a, D.affine(logtau, 0.5 * MN, -0.5 * MN * LOG2PI))

    def bound_col(self):
        """E[log p(X)] - E[log q(X)] from the statistics (no pass over N)."""
        col, N, K = self.col, self.Ng, self.K
        fz = col._fused
        u_par = col.moments_from_parents()
        cgf = D.asarray(col._distribution.compute_cgf_from_parents(*u_par))     # per-plate, shared
        _, Sxx, sx = self.stats()
        Lam, logdet, cov = fz["Lam"], fz["logdet"], fz["cov"]
        # the prior is evaluated NOW (its parents may have moved since q(X) was formed); q's own
        # natural parameters are the ones stored by update_col
        phi_p = col._canonical_phi(col._distribution.compute_phi_from_parents(*u_par))
        if not (all(n == 1 for n in phi_p[0].shape[:-1]) and all(n == 1 for n in phi_p[1].shape[:-2])):
            return self._orig["col.lb"]()
        SXX = D.axpby(float(N), cov.reshape((K, K)), 1.0, Sxx)                  # sum_n <xx^T>
        trLS = D.sum_product([Lam, Sxx], [["i", "j"], ["i", "j"]], [])           # sum_n x^T Lam x
        # sum_n (phi0p - phi0q_n).x_n = phi0p.s_x - tr(Lam S_xx)
        t0 = D.sub(D.sum_product([phi_p[0].reshape((K,)), sx], [["k"], ["k"]], []), trLS)
        # sum_n (phi1p - phi1q):<xx^T>_n,  phi1q = -Lam/2
        dphi1 = D.axpby(1.0, phi_p[1].reshape((K, K)), 0.5, Lam)
        t1 = D.sum_product([dphi1, SXX], [["i", "j"], ["i", "j"]], [])
        # -sum_n g_n = 1/2 tr(Lam S_xx) - N/2 logdet
        mg = D.axpby(0.5, trLS, -0.5 * N, logdet)
        ncgf = D.mul(D.reduce_to_shape(cgf, (), from_shape=col.plates), float(self.Ng) / float(self.N))
        return D.add(D.add(D.add(t0, t1), mg), ncgf)


    # ---- device-resident VB loop (csrc/pca_vb.cu) ----------------------------------------------------
    @staticmethod
    def _const_vec(node, K):
        """Value of a Constant parent as a length-K host vector, or None if it varies over plates."""
        if not isinstance(node, Constant) or node.value is None:
            return None
        v = np.asarray(node.value, dtype=np.float64)
        if v.size == 1:
            return np.full(K, float(v.reshape(())))
        if v.size == K and v.shape[-1] == K:
            return v.reshape(K).copy()
        return None

    def _resident_hyper(self):
        """Host constants of the model, or None when some prior is not a plain constant."""
        K, col, row, tau = self.K, self.col, self.row, self.tau
        h = dict(mux=self._const_vec(col.parents[0], K), ax=self._const_vec(col.parents[1], K),
                 muc=self._const_vec(row.parents[0], K))
        alpha = row.parents[1]
        if isinstance(alpha, Gamma):
            if tuple(alpha.plates) != (K,) or len(alpha.children) != 1 or alpha.observed is not False:
                return None
            h["a0"] = self._const_vec(alpha.parents[0], K)
            h["b0"] = self._const_vec(alpha.parents[1], K)
        else:
            h["a0"] = h["b0"] = np.ones(K)
            h["alpha_const"] = self._const_vec(alpha, K)
            if h["alpha_const"] is None:
                return None
        if isinstance(tau, Gamma):
            if len(tau.children) != 1 or tau.observed is not False:
                return None
            ta0, tb0 = self._const_vec(tau.parents[0], 1), self._const_vec(tau.parents[1], 1)
            if ta0 is None or tb0 is None:
                return None
            h["ta0"], h["tb0"] = ta0[:1], tb0[:1]
        else:
            h["ta0"] = h["tb0"] = np.ones(1)
            tc = self._const_vec(tau, 1)
            if tc is None:
                return None
            h["tau_const"] = tc[:1]
        if any(v is None for v in h.values()):
            return None
        if np.any(h["ax"] <= 0):
            return None
        return h

    def resident_program(self, vb, nodes):
        """Opcode list for ONE iteration of ``VB.update(*nodes)`` when the whole sweep can stay on the
        device (this model, constant hyper-priors, every latent node updated exactly once), else None."""
        if not self._valid() or self.masked():
            return None
        col, row, tau, Y = self.col, self.row, self.tau, self.Y
        alpha = row.parents[1]
        # everything _resident_hyper() decides on: parent identities, observed state and fan-out of the hyper nodes
        # (re-read on every call: a node may be observed, re-parented or given another child between two updates)
        hk = tuple(id(pp) for n in (col, row, alpha, tau) for pp in getattr(n, "parents", ())) + (id(alpha), id(tau)) \
            + tuple((o if (o is True or o is False or o is None) else "array", len(getattr(n, "children", ())))
                    for n in (alpha, tau) for o in (getattr(n, "observed", None),))
        key = (tuple(map(id, nodes)), tuple(map(id, vb.model)), hk)
        cached = getattr(self, "_prog_cache", None)
        if cached is not None and cached[0] == key:
            return cached[1]
        self._prog_cache = None
        latent = {col, row}
        if isinstance(alpha, Gamma):
            latent.add(alpha)
        if isinstance(tau, Gamma):
            latent.add(tau)
        model = [n for n in vb.model if n is not self.F]
        if set(model) != latent | {Y} or len(model) != len(latent) + 1:
            return None
        order = [vb[n] for n in nodes]
        order = [n for n in order if n is not Y and n is not self.F]
        if set(order) != latent or len(order) != len(latent):
            return None
        if getattr(self, "_hyper_key", None) != hk:
            self._res_cache = None
            self._hyper = self._resident_hyper()
            self._hyper_key = hk
            self._hsig = None if self._hyper is None else \
                tuple(np.asarray(v, dtype=np.float64).tobytes() for _, v in sorted(self._hyper.items()))
        if self._hyper is None:
            return None
        V = _bpk.VBOP
        ops = []
        for n in order:
            if n is col:
                ops += [V["XPRE"], V["XSWEEP"], V["STATS"], V["SXXT"]]
            elif n is row:
                ops.append(V["ROW"])
            elif n is alpha:
                ops.append(V["ALPHA"])
            else:
                ops.append(V["TAU"])
        ops.append(V["BOUND"])
        self._prog_cache = (key, (ops, order))
        return ops, order

    def _resident_enter(self, order, lprev):
        """Build the state vector from the node graph (a handful of small D2H reads, once per update call)."""
        be = _bpk.get()
        M, N, K, h = self.M, self.N, self.K, self._hyper
        lay, total = be.pca_vb_layout(M, K)
        st = np.zeros(total)

        def put(name, v):
            o, n = lay[name]
            st[o:o + n] = np.asarray(v, dtype=np.float64).reshape(-1)
        for k in ("mux", "ax", "muc", "a0", "b0", "ta0", "tb0"):
            put(k, h[k])
        put("ng", float(self.Ng))
        put("sumsq", self._sumsq_y().numpy())
        put("lprev", lprev)
        W = self._mean(self.row)
        if W.shape[0] != M:
            W = self.row.u[0].broadcast_to((M, 1, K)).reshape((M, K))
        put("w", W.numpy())
        put("sww", self._sum_second_moment(self.row, M).numpy())
        alpha = self.row.parents[1]
        if isinstance(alpha, Gamma):
            put("al_u0", np.broadcast_to(alpha.u[0].numpy(), (K,)))
            put("al_u1", np.broadcast_to(alpha.u[1].numpy(), (K,)))
        else:
            put("al_u0", h["alpha_const"])
            with np.errstate(divide="ignore"):
                put("al_u1", np.log(h["alpha_const"]))
        if isinstance(self.tau, Gamma):
            put("tau_u0", self.tau.u[0].numpy())
            put("tau_u1", self.tau.u[1].numpy())
        else:
            put("tau_u0", h["tau_const"])
            put("tau_u1", np.log(h["tau_const"]))
        if order[0] is not self.col:
            # somebody consumes the statistics of the CURRENT q(X) before the first sweep
            Syx, Sxx, sx = self.stats()
            put("stats", self._stats[1].numpy())
            put("sxxt", self._sum_xx().numpy())
        state = DArray.from_numpy(st)
        return state, lay

    def _resident_state(self, order, lprev, need_prev=False):
        """Device state vector + X buffer for a resident run: re-used straight after a resident run when nothing
        touched the nodes in between (the state and X of that run are still exact), else rebuilt from the nodes."""
        alpha_n = self.row.parents[1]
        watched = [self.Y, self.col, self.row] + [n for n in (alpha_n, self.tau) if isinstance(n, Gamma)]
        hsig = self._hsig
        cache = getattr(self, "_res_cache", None)
        if cache is not None and cache["versions"] == [n._version for n in watched] and cache["hsig"] == hsig \
                and cache["nodes"] == [id(n) for n in watched] and not (need_prev and np.isnan(lprev)):
            state, lay, X = cache["state"], cache["lay"], cache["X"]
        else:
            cache = None
            state, lay = self._resident_enter(order, lprev)
            X = DArray.empty((1, self.N, self.K))
        self._res_cache = None
        return state, lay, X, cache, watched, hsig

    @staticmethod
    def _raise_ctrl_errors(err):
        if err & 1:
            raise _bpk.NotPositiveDefinite("Matrix not positive definite")
        if err & 2:
            raise ValueError("Natural parameters should be positive")
        if err & 4:
            raise _bpk.BpkError(_bpk.ENCCL, "peer-memory exchange timed out (a rank stopped responding)")
        if err & 8:
            raise _bpk.BpkError(_bpk.ECUDA, "grid barrier of the fused sweep kernel timed out (is another process "
                                            "using this GPU? the persistent grid must be fully resident)")

    def sweep_resident(self, vb, program):
        """ONE sweep of ``VB.update`` — every node update of the user's order, no lower bound — as one fused launch.
        Used when a callback (e.g. the rotation of transformations.py) must run between the node updates and the
        bound evaluation (vmp.py:702-713): the heavy part of the iteration stays on the device path, the callback
        sees and may modify the node objects, and the bound is then evaluated from the plate-summed statistics."""
        be = _bpk.get()
        ops, order = program
        ops = [o for o in ops if o != _bpk.VBOP["BOUND"]]
        M, N, K = self.M, self.N, self.K
        self._agree_on_loop_mode(be)
        state, lay, X, cache, watched, hsig = self._resident_state(order, np.nan)
        alpha = self.row.parents[1]
        ctrl = DArray.zeros((2,))
        Lh = DArray.empty((1, 6))
        be.pca_vb_run(self._Yd().ptr, M, N, K, X.ptr, state.ptr, ops, 1, isinstance(alpha, Gamma),
                      isinstance(self.tau, Gamma), -1.0, Lh.ptr, 0, ctrl.ptr)
        c = ctrl.numpy().view(np.int32)
        self._raise_ctrl_errors(int(c[2]))
        self.fused_calls += 1
        if cache is not None and state is cache["state"]:
            self._resident_republish(X)
        else:
            self._resident_publish(state, lay, X, order)
        self._res_cache = dict(state=state, lay=lay, X=X, hsig=hsig, nodes=[id(n) for n in watched],
                               versions=[n._version for n in watched])

    def _agree_on_loop_mode(self, be):
        if self.world > 1 and hasattr(be, "pca_vb_set_mode"):
            # the persistent multi-sweep launch needs 16-byte aligned rows (even N, aligned Y); every rank must
            # take the same route through the in-kernel exchange, so agree on it once per data buffer
            Yp = self._Yd().ptr
            key = (Yp, self.N)
            if getattr(self, "_loop_key", None) != key:
                mine = 1.0 if (self.N % 2 == 0 and Yp % 16 == 0) else 0.0
                self._loop_all = bool(np.min(parallel.allgather_scalar(mine)) > 0.5)
                self._loop_key = key
            be.pca_vb_set_mode(not self._loop_all)

    def run_resident(self, vb, program, repeat, tol, verbose):
        """``VB.update`` for this model without leaving the device between sweeps."""
        be = _bpk.get()
        ops, order = program
        M, N, K = self.M, self.N, self.K
        import time
        import warnings
        self._agree_on_loop_mode(be)
        check = not vb.ignore_bound_checks
        tol_dev = (vb.tol if tol is None else tol) if check else -1.0
        lprev = np.nan
        if check and not vb.annealing_changed and vb.iter > 0:
            lprev = vb.L[vb.iter - 1]
        state, lay, X, cache, watched, hsig = self._resident_state(order, lprev, need_prev=check and vb.iter > 0)
        fast = M <= 64 and K <= 16
        alpha = self.row.parents[1]
        has_alpha, has_tau = isinstance(alpha, Gamma), isinstance(self.tau, Gamma)
        terms_of = {self.Y: 0, self.col: 1, self.row: 2, alpha: 3, self.tau: 4}
        Yd = self._Yd()
        done, converged = 0, False
        t_enter = time.perf_counter()
        host = dict(pre_us=0.0, call_us=0.0, wait_us=0.0, post_us=0.0, launches=0)
        while (repeat is None or done < repeat) and not converged:
            left = 50 if repeat is None else repeat - done
            chunk = 1 if (verbose or not fast) else min(left, 50)
            # bound rows and the control words share one buffer (kept across calls): one read-back per chunk
            buf = getattr(self, "_ctl_buf", None)
            if buf is None or buf.size != chunk * 6 + 2:
                buf = self._ctl_buf = DArray.empty((chunk * 6 + 2,))
            ctrl_ptr = buf.ptr + chunk * 48
            be.memset(ctrl_ptr, 0, 16)            # int[4]: iterations done, stop, error bits
            if self.kernel_timers is not None:
                ids = self.kernel_timers[self._timer_pos:self._timer_pos + chunk]
                be.pca_vb_set_timers(ids)
            t0 = time.time()
            tc0 = time.perf_counter()
            be.pca_vb_run(Yd.ptr, M, N, K, X.ptr, state.ptr, ops, chunk, has_alpha, has_tau, tol_dev,
                          buf.ptr, chunk, ctrl_ptr)
            tc1 = time.perf_counter()
            hbuf = buf.numpy()                    # blocks until the chunk has run
            tc2 = time.perf_counter()
            c = hbuf[chunk * 6:].view(np.int32)
            host["pre_us"] += 1e6 * (tc0 - t_enter)
            host["call_us"] += 1e6 * (tc1 - tc0)
            host["wait_us"] += 1e6 * (tc2 - tc1)
            host["launches"] += 1
            if self.kernel_timers is not None:
                used = be.pca_vb_timers_used()
                # one timer may bracket several sweeps (the whole chunk is one launch when it can be)
                self.kernel_timer_log.append((self.kernel_timers[self._timer_pos:self._timer_pos + used], None))
                self._timer_pos += used
                be.pca_vb_set_timers([])
            dt = (time.time() - t0)
            n_it, stop, err = int(c[0]), int(c[1]), int(c[2])
            if self.kernel_timers is not None and self.kernel_timer_log and self.kernel_timer_log[-1][1] is None:
                self.kernel_timer_log[-1] = (self.kernel_timer_log[-1][0], n_it)
            self._raise_ctrl_errors(err)
            vb._record_resident_iterations(hbuf[:chunk * 6].reshape(chunk, 6)[:n_it], terms_of, dt, stop, check, verbose)
            done += n_it
            self.fused_calls += n_it
            converged = bool(stop)
            t_enter = time.perf_counter()
            host["post_us"] += 1e6 * (t_enter - tc2)
            if n_it == 0:
                break
        if done > 0:
            if cache is not None and state is cache["state"]:
                self._resident_republish(X)
            else:
                self._resident_publish(state, lay, X, order)
            self._res_cache = dict(state=state, lay=lay, X=X, hsig=hsig, nodes=[id(n) for n in watched],
                                   versions=[n._version for n in watched])
        host["post_us"] += 1e6 * (time.perf_counter() - t_enter)
        self.host_timing = host                   # where the host spent its time around the launches (bench.py reports it)
        return converged

    def _resident_republish(self, X):
        """Same state vector and X buffer as the previous resident run: the node objects already view them;
        only what caches derived values has to be renewed (no kernel launches)."""
        col, row, tau = self.col, self.row, self.tau
        N, K = self.N, self.K
        fz = col._fused
        Lam, logdet, cov = fz["Lam"], fz["logdet"], fz["cov"]

        def phi0_fn():
            return D.sum_product([Lam, X], [["i", "j"], ["o", "n", "j"]], ["o", "n", "i"])

        def g_fn():
            q = D.sum_product([X, Lam, X], [["o", "n", "i"], ["i", "j"], ["o", "n", "j"]], ["o", "n"])
            return D.axpby(-0.5, q, 0.5, logdet)
        col.phi = [LazyArray((1, N, K), phi0_fn), col.phi[1]]
        col.u = [X, FactoredSecondMoment(X, cov, (K,))]
        col.g = LazyArray((1, N), g_fn)
        col._version += 1
        self._stats = (col._version, self._stats[1])
        row.u = [row.u[0], FactoredSecondMoment(row.u[0], row.u[1].cov, (K,))]
        row._version += 1
        alpha = row.parents[1]
        if isinstance(alpha, Gamma):
            alpha._version += 1
        if isinstance(tau, Gamma):
            tau._version += 1
        self._e2 = None

    def _resident_publish(self, state, lay, X, order):
        """Point the node objects at the device state the loop left behind."""
        M, N, K = self.M, self.N, self.K
        col, row, tau = self.col, self.row, self.tau

        def view(name, shape):
            o, n = lay[name]
            return state.slice_axis(0, o, o + n).reshape(shape)
        # X
        Lam = view("lamx", (K, K))
        cov = view("covx", (1, 1, K, K))
        logdet = view("logdetx", ())
        phi1 = view("phi1x", (1, 1, K, K))

        def phi0_fn():
            return D.sum_product([Lam, X], [["i", "j"], ["o", "n", "j"]], ["o", "n", "i"])

        def g_fn():
            q = D.sum_product([X, Lam, X], [["o", "n", "i"], ["i", "j"], ["o", "n", "j"]], ["o", "n"])
            return D.axpby(-0.5, q, 0.5, logdet)
        col.phi = [LazyArray((1, N, K), phi0_fn), phi1]
        col.u = [X, FactoredSecondMoment(X, cov, (K,))]
        col.g = LazyArray((1, N), g_fn)
        col._version += 1
        col._fused = dict(Lam=Lam, logdet=logdet, cov=cov, phi_p=None)     # bound_col re-evaluates the prior itself
        self._stats = (col._version, view("stats", (M * K + K * K + K,)))
        # C
        W = view("w", (M, 1, K))
        row.phi = [view("phi0c", (M, 1, K)), view("phi1c", (1, 1, K, K))]
        row.u = [W, FactoredSecondMoment(W, view("covc", (1, 1, K, K)), (K,))]
        row.g = view("gc", (M, 1))
        row._version += 1
        alpha = row.parents[1]
        if isinstance(alpha, Gamma):
            alpha.phi = [view("al_phi0", (K,)), view("al_phi1", (K,))]
            alpha.u = [view("al_u0", (K,)), view("al_u1", (K,))]
            alpha.g = view("al_g", (K,))
            alpha._version += 1
        if isinstance(tau, Gamma):
            shp = tuple(tau.plates)
            tau.phi = [view("tau_phi0", shp), view("tau_phi1", shp)]
            tau.u = [view("tau_u0", shp), view("tau_u1", shp)]
            tau.g = view("tau_g", shp)
            tau._version += 1
        self._e2 = None


class GaussianMixturePlan:
    """Y = Mixture(Z, Gaussian, mu, Lambda) with Y fully observed (gmm.rst:71-98).

    Z.update()                  -> bpk_gmm_sweep: one pass over y writes the responsibilities
                                   and accumulates  R_k = sum_n p_nk, s1_k = sum_n p_nk y_n,
                                   S2_k = sum_n p_nk y_n y_n^T  and  sum_n logsumexp_n
    Y.message_to_parent(mu)     -> [<Lam_k> s1_k, -1/2 <Lam_k> R_k]                 (mixture.py:108-160)
    Y.message_to_parent(Lambda) -> [-1/2 (S2 - s1 mu^T - mu s1^T + R <mu mu^T>), 1/2 R]
    Z.message_to_parent(alpha)  -> [R]                                              (multinomial.py:83-90)
    Y / Z lower bounds          -> from the statistics                              (expfamily.py:400-480)
    """

    def __init__(self, Y, Z, mu, Lam):
        self.Y, self.Z, self.mu, self.Lam = Y, Z, mu, Lam
        self.N = Y.plates[0]
        self.D = mu.dims[0][0]
        self.K = mu.plates[0]
        self.world = parallel.world()
        self.Ng = int(round(float(np.sum(parallel.allgather_scalar(self.N))))) if self.world > 1 else self.N
        self._stats = None           # (Z._version, Y._version, DArray)
        self._orig = {}
        self.fused_calls = 0
        self.kernel_timers = None
        self._timer_pos = 0
        self._install()

    def _install(self):
        plan, o = self, self._orig
        o["Z.update"] = self.Z.update
        o["Z.lb"] = self.Z.lower_bound_contribution
        o["Z.msg"] = self.Z.message_to_parent
        o["Y.msg"] = self.Y.message_to_parent
        o["Y.lb"] = self.Y.lower_bound_contribution

        def Z_update(node, annealing=1.0):
            if annealing == 1.0 and plan.valid():
                return plan.update_Z()
            return o["Z.update"](annealing) if annealing != 1.0 else o["Z.update"]()

        def Z_lb(node, ignore_masked=True):
            if not ignore_masked:
                return o["Z.lb"](ignore_masked=False)
            return plan.bound_Z() if plan.valid() and getattr(plan.Z, "_fused", None) is not None else o["Z.lb"]()

        def Z_msg(node, index):
            if index == 0 and plan.valid() and mask_is_full(plan.Z.mask):
                return plan.message_to_alpha()
            return o["Z.msg"](index)

        def Y_msg(node, index):
            if index in (1, 2) and plan.valid():
                return plan.message_to_mu() if index == 1 else plan.message_to_Lambda()
            return o["Y.msg"](index)

        def Y_lb(node, ignore_masked=True):
            if not ignore_masked:
                return o["Y.lb"](ignore_masked=False)
            return plan.bound_Y() if plan.valid() else o["Y.lb"]()

        self.Z.update = types.MethodType(Z_update, self.Z)
        self.Z.lower_bound_contribution = types.MethodType(Z_lb, self.Z)
        self.Z.message_to_parent = types.MethodType(Z_msg, self.Z)
        self.Y.message_to_parent = types.MethodType(Y_msg, self.Y)
        self.Y.lower_bound_contribution = types.MethodType(Y_lb, self.Y)

    def valid(self):
        Y, Z = self.Y, self.Z
        ok = (Y.observed is True) and mask_is_full(Y.mask) and Z.observed is False \
            and all(getattr(n, "annealing", 1.0) == 1.0 for n in (Y, Z, self.mu, self.Lam)) \
            and not any(getattr(n, "plates_multiplier", ()) for n in (Y, Z, self.mu, self.Lam))
        if not ok and self.world > 1:
            raise NotImplementedError("plate sharding is only supported on the fused path")
        return ok

    # ---- cluster parameters as the kernel wants them --------------------------------------------
    def _params(self):
        """g_k (K), h_k = <Lam_k mu_k> (K,D), <Lam_k> (K,D,D) from the mixed Gaussian's protocol."""
        raw = self.Y._distribution.raw
        u_mu, u_L = self.mu.get_moments(), self.Lam.get_moments()
        phi = raw.compute_phi_from_parents(u_mu, u_L)
        g = D.asarray(raw.compute_cgf_from_parents(u_mu, u_L))
        K, Dm = self.K, self.D
        h = D.asarray(phi[0]).broadcast_to((K, Dm)).contiguous()
        Lm = D.asarray(u_L[0]).broadcast_to((K, Dm, Dm)).contiguous()
        return g.broadcast_to((K,)).contiguous(), h, Lm

    def _Yd(self):
        return self.Y.u[0].contiguous()

    def stats(self):
        key = (self.Z._version, self.Y._version)
        if self._stats is None or self._stats[0] != key:
            N, Dm, K = self.N, self.D, self.K
            st = DArray.zeros((K + K * Dm + K * Dm * Dm + 1,))
            P = self.Z.u[0].broadcast_to((N, K)).contiguous()
            _bpk.get().gmm_stats(self._Yd().ptr, N, Dm, K, P.ptr, st.ptr)
            parallel.allreduce_sum(st)
            self._stats = (key, st)
        return self._split(self._stats[1])

    def _split(self, st):
        K, Dm = self.K, self.D
        a, b, c = K, K + K * Dm, K + K * Dm + K * Dm * Dm
        return (st.slice_axis(0, 0, a), st.slice_axis(0, a, b).reshape((K, Dm)),
                st.slice_axis(0, b, c).reshape((K, Dm, Dm)), st.slice_axis(0, c, c + 1).reshape(()))

    # ---- fused operations --------------------------------------------------------------------------
    def update_Z(self):
        Z, N, Dm, K = self.Z, self.N, self.D, self.K
        be = _bpk.get()
        u_par = Z.moments_from_parents()
        logpi = D.asarray(Z._distribution.compute_phi_from_parents(*u_par)[0])
        if logpi.size != K:
            return self._orig["Z.update"]()           # per-sample prior: generic path
        logpi = logpi.reshape((K,)).contiguous()
        g, h, Lm = self._params()
        P = DArray.empty((N, K))
        gz = DArray.empty((N,))
        st = DArray.zeros((K + K * Dm + K * Dm * Dm + 1,))
        Yd = self._Yd()
        if self.kernel_timers is not None and self._timer_pos < len(self.kernel_timers):
            tid = self.kernel_timers[self._timer_pos]
            self._timer_pos += 1
            be.timer_record(tid, 0)
            be.gmm_sweep(Yd.ptr, N, Dm, K, g.ptr, h.ptr, Lm.ptr, logpi.ptr, P.ptr, gz.ptr, st.ptr)
            be.timer_record(tid, 1)
        else:
            be.gmm_sweep(Yd.ptr, N, Dm, K, g.ptr, h.ptr, Lm.ptr, logpi.ptr, P.ptr, gz.ptr, st.ptr)
        parallel.allreduce_sum(st)
        self.fused_calls += 1
        Y = self.Y

        def phi_fn():
            m = plan_self._orig["Y.msg"](0)[0]
            return D.add(m, logpi)
        plan_self = self
        Z.phi = [LazyArray((N, K), phi_fn)]
        Z.u = [P]
        Z.g = gz
        Z._version += 1
        self._stats = ((Z._version, Y._version), st)
        # sum_nk p_nk * (message from Y) with the parameters q(Z) was built from: needed by Z's bound
        Z._fused = dict(logpi=logpi, T=self._T(params=(g, h, Lm)))

    def message_to_alpha(self):
        R, _, _, _ = self.stats()
        self.fused_calls += 1
        par = self.Z.parents[0]
        return [R.reshape((1,) * len(par.plates) + (self.K,)) if len(par.plates) else R]

    def message_to_mu(self):
        R, s1, _, _ = self.stats()
        Lm = D.asarray(self.Lam.get_moments()[0]).broadcast_to((self.K, self.D, self.D))
        m0 = D.sum_product([Lm, s1], [["k", "i", "j"], ["k", "j"]], ["k", "i"])
        m1 = D.mul(D.mul(Lm, R.reshape((self.K, 1, 1))), -0.5)
        self.fused_calls += 1
        return [m0, m1]

    def message_to_Lambda(self):
        R, s1, S2, _ = self.stats()
        u_mu = self.mu.get_moments()
        K, Dm = self.K, self.D
        mu = D.asarray(u_mu[0]).broadcast_to((K, Dm))
        mumu = D.asarray(u_mu[1]).broadcast_to((K, Dm, Dm))
        sm = D.mul(s1.reshape((K, Dm, 1)), mu.reshape((K, 1, Dm)))            # s1 mu^T
        ms = D.mul(mu.reshape((K, Dm, 1)), s1.reshape((K, 1, Dm)))            # mu s1^T
        t = D.add(D.sub(D.sub(S2, sm), ms), D.mul(mumu, R.reshape((K, 1, 1))))
        self.fused_calls += 1
        return [D.mul(t, -0.5), D.mul(R, 0.5)]

    def _T(self, params=None):
        """sum_nk p_nk (g_k + h_k.y_n - 1/2 y_n^T Lam_k y_n) from the statistics."""
        R, s1, S2, _ = self.stats()
        g, h, Lm = params if params is not None else self._params()
        a = D.sum_product([g, R], [["k"], ["k"]], [])
        b = D.sum_product([h, s1], [["k", "i"], ["k", "i"]], [])
        c = D.sum_product([Lm, S2], [["k", "i", "j"], ["k", "i", "j"]], [], scale=-0.5)
        return D.add(D.add(a, b), c)

    def bound_Y(self):
        """E[log p(Y|Z,mu,Lambda)] = T - N D/2 log 2pi."""
        return D.affine(self._T(), 1.0, -0.5 * self.D * LOG2PI * self.Ng)

    def bound_Z(self):
        """E[log p(Z|pi)] - E[log q(Z)] = sum_n logsumexp_n - T_q + sum_k (<log pi_k> - logpi_q,k) R_k,
        where T_q and logpi_q are the message sum and prior q(Z) was built from (expfamily.py:400-480)."""
        R, _, _, lse = self.stats()
        fz = self.Z._fused
        u_par = self.Z.moments_from_parents()
        logpi = D.asarray(self.Z._distribution.compute_phi_from_parents(*u_par)[0]).reshape((self.K,))
        dpi = D.sum_product([D.sub(logpi, fz["logpi"]), R], [["k"], ["k"]], [])
        return D.add(D.sub(lse, fz["T"]), dpi)


    # ---- device-resident loop (csrc/gmm_vb.cu): whole sweeps without returning to Python -----------------
    def _hyper_nodes(self):
        return self.mu, self.Lam, self.Z.parents[0]

    def _resident_hyper(self):
        """Prior natural parameters and log-normalisers of mu, Lambda, alpha as host arrays, or None when one of
        them does not hang off plain constants."""
        from .dirichlet import Dirichlet
        K, Dm = self.K, self.D
        mu, Lam, alpha = self._hyper_nodes()
        if not isinstance(alpha, Dirichlet) or tuple(alpha.plates) != () or tuple(alpha.dims[0]) != (K,) \
                or len(alpha.children) != 1:
            return None
        h = {}
        for key, n, shapes, gshape in (("m", mu, [(K, Dm), (K, Dm, Dm)], (K,)), ("l", Lam, [(K, Dm, Dm), (K,)], (K,)),
                                      ("a", alpha, [(K,)], (1,))):
            if n.observed is not False or not all(isinstance(pp, Constant) for pp in n.parents) \
                    or getattr(n, "annealing", 1.0) != 1.0 or not mask_is_full(n.mask):
                return None
            u_par = n.moments_from_parents()
            phi = n._canonical_phi(n._distribution.compute_phi_from_parents(*u_par))
            g = D.asarray(n._distribution.compute_cgf_from_parents(*u_par))
            for i, (ph, shp) in enumerate(zip(phi, shapes)):
                h["p%s%d" % (key, i)] = np.broadcast_to(dense(ph).numpy(), shp).astype(np.float64).ravel()
            h["gp" + key] = np.broadcast_to(g.numpy(), gshape).astype(np.float64).ravel()
        return h

    def resident_program(self, vb, nodes):
        """Opcode list for ONE iteration of ``VB.update(*nodes)`` when the whole sweep can stay on the device
        (this model, constant hyper-priors, every latent node updated exactly once), else None."""
        be = _bpk.get()
        if not hasattr(be, "gmm_vb_run") or not self.valid():
            return None
        mu, Lam, alpha = self._hyper_nodes()
        Z, Y = self.Z, self.Y
        latent = [Z, mu, Lam, alpha]
        if len(vb.model) != 5 or any(n not in vb.model for n in latent + [Y]):
            return None
        try:
            order = [vb[n] for n in nodes]
        except Exception:
            return None
        order = [n for n in order if n is not Y]
        if len(order) != 4 or any(n not in order for n in latent):
            return None
        hk = tuple(id(pp) for n in (mu, Lam, alpha) for pp in n.parents) + (id(alpha),) \
            + tuple((repr(n.observed) if not isinstance(n.observed, np.ndarray) else "array", len(n.children))
                    for n in (mu, Lam, alpha))
        if getattr(self, "_hyper_key", None) != hk:
            self._res_cache = None
            self._hyper = self._resident_hyper()
            self._hyper_key = hk
            self._hsig = None if self._hyper is None else tuple(v.tobytes() for _, v in sorted(self._hyper.items()))
        if self._hyper is None:
            return None
        for n in (mu, Lam, alpha):
            if n.phi is None or n.u is None or n.g is None or any(x is None for x in list(n.phi) + list(n.u)):
                return None
        G = _bpk.GMMOP
        code = {id(Z): G["Z"], id(mu): G["MU"], id(Lam): G["LAMBDA"], id(alpha): G["ALPHA"]}
        return [code[id(n)] for n in order] + [G["BOUND"]], order

    def _resident_enter(self, order, lprev):
        be = _bpk.get()
        K, Dm, h = self.K, self.D, self._hyper
        mu, Lam, alpha = self._hyper_nodes()
        lay, total = be.gmm_vb_layout(Dm, K)
        st = np.zeros(total)

        def put(name, v, shape=None):
            v = np.asarray(dense(v).numpy() if hasattr(dense(v), "numpy") else v, dtype=np.float64)
            if shape is not None:
                v = np.broadcast_to(v, shape)
            v = v.reshape(-1)
            o = lay[name][0]
            st[o:o + v.size] = v
        put("pm0", h["pm0"]); put("pm1", h["pm1"]); put("gpm", h["gpm"])
        put("pl0", h["pl0"]); put("pl1", h["pl1"]); put("gpl", h["gpl"])
        put("pa", h["pa0"]); put("gpa", h["gpa"])
        put("ng", float(self.Ng))
        put("lprev", lprev)
        put("mu_phi0", mu.phi[0], (K, Dm)); put("mu_phi1", mu.phi[1], (K, Dm, Dm))
        put("mu_u0", mu.u[0], (K, Dm)); put("mu_u1", mu.u[1], (K, Dm, Dm)); put("mu_g", mu.g, (K,))
        cov = getattr(mu.u[1], "cov", None)
        if cov is not None:
            put("mu_cov", cov, (K, Dm, Dm))
        put("lam_phi0", Lam.phi[0], (K, Dm, Dm)); put("lam_phi1", Lam.phi[1], (K,))
        put("lam_u0", Lam.u[0], (K, Dm, Dm)); put("lam_u1", Lam.u[1], (K,)); put("lam_g", Lam.g, (K,))
        put("al_phi", alpha.phi[0], (K,)); put("al_u", alpha.u[0], (K,)); put("al_g", alpha.g, (1,))
        if order[0] is not self.Z:
            # somebody consumes the statistics of the CURRENT q(Z) before the first sweep
            self.stats()
            put("stats", self._stats[1])
            fz = getattr(self.Z, "_fused", None)
            if fz is not None:
                put("z_logpi", fz["logpi"], (K,))
                put("z_t", fz["T"], (1,))
        return DArray.from_numpy(st), lay

    def _resident_state(self, order, lprev, need_prev=False):
        mu, Lam, alpha = self._hyper_nodes()
        watched = [self.Y, self.Z, mu, Lam, alpha]
        cache = getattr(self, "_res_cache", None)
        if cache is not None and cache["versions"] == [n._version for n in watched] and cache["hsig"] == self._hsig \
                and cache["nodes"] == [id(n) for n in watched] and not (need_prev and np.isnan(lprev)):
            state, lay, P, gz = cache["state"], cache["lay"], cache["P"], cache["gz"]
        else:
            cache = None
            state, lay = self._resident_enter(order, lprev)
            P, gz = DArray.empty((self.N, self.K)), DArray.empty((self.N,))
        self._res_cache = None
        return state, lay, P, gz, cache, watched

    @staticmethod
    def _raise_ctrl_errors(err):
        if err & 1:
            raise _bpk.NotPositiveDefinite("Matrix not positive definite")
        if err & 2:
            raise ValueError("Natural parameters should be positive")

    def _resident_publish(self, state, lay, P, gz):
        """Point the node objects at the device state the loop left behind."""
        K, Dm, N = self.K, self.D, self.N
        mu, Lam, alpha = self._hyper_nodes()
        Z, Y = self.Z, self.Y

        def view(name, shape):
            o = lay[name][0]
            n = int(np.prod(shape, dtype=np.int64)) if shape else 1
            return state.slice_axis(0, o, o + n).reshape(shape)
        mu.phi = [view("mu_phi0", (K, Dm)), view("mu_phi1", (K, Dm, Dm))]
        u0 = view("mu_u0", (K, Dm))
        second = FactoredSecondMoment(u0, view("mu_cov", (K, Dm, Dm)), (Dm,))
        second._dense = view("mu_u1", (K, Dm, Dm))
        mu.u = [u0, second]
        mu.g = view("mu_g", (K,))
        mu._version += 1
        Lam.phi = [view("lam_phi0", (K, Dm, Dm)), view("lam_phi1", (K,))]
        Lam.u = [view("lam_u0", (K, Dm, Dm)), view("lam_u1", (K,))]
        Lam.g = view("lam_g", (K,))
        Lam._version += 1
        alpha.phi = [view("al_phi", (K,))]
        alpha.u = [view("al_u", (K,))]
        alpha.g = view("al_g", ())
        alpha._version += 1
        logpi = view("z_logpi", (K,))
        plan = self

        def phi_fn():
            return D.add(plan._orig["Y.msg"](0)[0], logpi)
        Z.phi = [LazyArray((N, K), phi_fn)]
        Z.u = [P]
        Z.g = gz
        Z._version += 1
        Z._fused = dict(logpi=logpi, T=view("z_t", ()))
        self._stats = ((Z._version, Y._version), view("stats", (K + K * Dm + K * Dm * Dm + 1,)))

    def _resident_finish(self, state, lay, P, gz, watched):
        self._resident_publish(state, lay, P, gz)
        self._res_cache = dict(state=state, lay=lay, P=P, gz=gz, hsig=self._hsig, nodes=[id(n) for n in watched],
                               versions=[n._version for n in watched])

    def sweep_resident(self, vb, program):
        """ONE sweep of ``VB.update`` — every node update of the user's order, no lower bound — without returning
        to Python between the nodes (a callback / progress bar wants the host before the bound, vmp.py:702-713)."""
        be = _bpk.get()
        ops, order = program
        ops = [o for o in ops if o != _bpk.GMMOP["BOUND"]]
        state, lay, P, gz, cache, watched = self._resident_state(order, np.nan)
        ctrl = DArray.zeros((2,))
        Lh = DArray.empty((1, 6))
        be.gmm_vb_run(self._Yd().ptr, self.N, self.D, self.K, P.ptr, gz.ptr, state.ptr, ops, 1, -1.0, Lh.ptr, 0, ctrl.ptr)
        c = ctrl.numpy().view(np.int32)
        self._raise_ctrl_errors(int(c[2]))
        self.fused_calls += 1
        self._resident_finish(state, lay, P, gz, watched)

    def run_resident(self, vb, program, repeat, tol, verbose):
        """``VB.update`` for this model without leaving the device between sweeps."""
        import time
        import warnings
        be = _bpk.get()
        ops, order = program
        mu, Lam, alpha = self._hyper_nodes()
        check = not vb.ignore_bound_checks
        tol_dev = (vb.tol if tol is None else tol) if check else -1.0
        lprev = np.nan
        if check and not vb.annealing_changed and vb.iter > 0:
            lprev = vb.L[vb.iter - 1]
        state, lay, P, gz, cache, watched = self._resident_state(order, lprev, need_prev=check and vb.iter > 0)
        if cache is not None:
            o = lay["lprev"][0]
            D.copy_into(state.slice_axis(0, o, o + 1), DArray.from_numpy(np.array([lprev])))
        terms_of = {self.Y: 0, self.Z: 1, mu: 2, Lam: 3, alpha: 4}
        Yd = self._Yd()
        done, converged = 0, False
        ctrl = DArray.zeros((2,))                 # 16 bytes = int[4]
        while (repeat is None or done < repeat) and not converged:
            left = 25 if repeat is None else repeat - done
            chunk = 1 if verbose else min(left, 25)
            Lh = DArray.empty((chunk, 6))
            be.memset(ctrl.ptr, 0, 16)
            ids = []
            if self.kernel_timers is not None:
                ids = self.kernel_timers[self._timer_pos:self._timer_pos + chunk]
                be.gmm_vb_set_timers(ids)
            t0 = time.time()
            be.gmm_vb_run(Yd.ptr, self.N, self.D, self.K, P.ptr, gz.ptr, state.ptr, ops, chunk, tol_dev, Lh.ptr, chunk,
                          ctrl.ptr)
            c = ctrl.numpy().view(np.int32)       # blocks until the chunk has run
            dt = time.time() - t0
            n_it, stop, err = int(c[0]), int(c[1]), int(c[2])
            if ids:
                self._timer_pos += min(len(ids), max(n_it, 0))
                be.gmm_vb_set_timers([])
            self._raise_ctrl_errors(err)
            vb._record_resident_iterations(Lh.numpy()[:n_it], terms_of, dt, stop, check, verbose)
            done += n_it
            self.fused_calls += n_it
            converged = bool(stop)
            if n_it == 0:
                break
        if done > 0:
            self._resident_finish(state, lay, P, gz, watched)
        return converged


def _attach_gmm(model, plans):
    from .mixture import Mixture
    from .gaussian import Gaussian
    from .wishart import Wishart
    from .categorical import Categorical
    for Y in model:
        if not isinstance(Y, Mixture) or Y.mixed_class is not Gaussian or Y.cluster_plate != -1:
            continue
        if getattr(Y, "_plan", None) is not None:
            plans.append(Y._plan)
            continue
        if len(Y.plates) != 1 or Y.children or len(Y.parents) != 3:
            continue
        Z, mu, Lam = Y.parents
        if not (isinstance(Z, Categorical) and isinstance(mu, Gaussian) and isinstance(Lam, Wishart)):
            continue
        if tuple(Z.plates) != tuple(Y.plates) or len(mu.plates) != 1 or tuple(mu.plates) != tuple(Lam.plates):
            continue
        if len(Z.children) != 1 or len(mu.children) != 1 or len(Lam.children) != 1:
            continue
        if mu.dims[0][0] > 16 or mu.plates[0] > 128:
            continue
        plan = GaussianMixturePlan(Y, Z, mu, Lam)
        Y._plan = plan
        plans.append(plan)


def attach(model):
    """Find fusable sub-graphs among ``model`` and install their plans."""
    plans = []
    for Y in model:
        if not isinstance(Y, GaussianARD) or len(Y.dims[0]) != 0 or len(Y.plates) != 2:
            continue
        if getattr(Y, "_plan", None) is not None:
            plans.append(Y._plan)
            continue
        F, tau = Y.parents
        if not isinstance(F, SumMultiply) or len(F.parents) != 2 or F.out_keys != []:
            continue
        if len(F.children) != 1 or Y.children:
            continue
        if not (isinstance(tau, (Gamma, Constant)) and _scalar_plates(tau)):
            continue
        A, B = F.parents
        if not (isinstance(A, GaussianARD) and isinstance(B, GaussianARD)):
            continue
        if len(A.dims[0]) != 1 or A.dims[0] != B.dims[0] or F.in_keys[0] != F.in_keys[1]:
            continue
        if len(A.children) != 1 or len(B.children) != 1:
            continue
        M, N = Y.plates
        pa, pb = tuple(A.plates), tuple(B.plates)
        if pa == (M, 1) and pb == (1, N):
            row, col, i_row, i_col = A, B, 0, 1
        elif pb == (M, 1) and pa == (1, N):
            row, col, i_row, i_col = B, A, 1, 0
        else:
            continue
        if A.dims[0][0] > _bpk.MAXDIM:
            continue
        plan = FactorModelPlan(Y, F, row, col, i_row, i_col, tau)
        Y._plan = plan
        plans.append(plan)
    _attach_gmm(model, plans)
    return plans
