"""Mixture node on device (replaces the array math of nodes/mixture.py:26-488).

``Y = Mixture(Z, NodeClass, *params, cluster_plate=-1)``: the parameters carry an extra
plate axis of length K (the clusters); Z holds the assignment probabilities.

    message to Z      L[n,k] = <g_k> + sum_i <phi_k,i> . u_n,i        (mixture.py:58-106)
    message to params p[n,k] x (the mixed distribution's message)      (mixture.py:108-160)
    phi / cgf         p-weighted averages over the clusters            (mixture.py:180-293)

The reference forms (N,K,D,D) temporaries for both; here every contraction is a
``sum_product`` (``bpk_sum_multiply``) over broadcast operands, and for the Gaussian case the
whole thing is served by the fused ``bpk_gmm_sweep`` plan (engine.plans.GaussianMixturePlan).
"""
import numpy as np

from .. import darray as D
from ..darray import DArray
from .categorical import categorical_constant
from .expfam import Distribution, ExponentialFamily
from .gaussian import dense
from .node import Node, broadcast_plates


def _move_cluster_last(a, cluster_axis):
    """View with the cluster axis (negative index among a's axes) moved to the end."""
    nd = a.ndim
    ax = cluster_axis + nd
    order = [i for i in range(nd) if i != ax] + [ax]
    return DArray(a.owner, a.ptr, [a.shape[i] for i in order], [a.strides[i] for i in order], a.dtype)


class MixtureDistribution(Distribution):

    def __init__(self, distribution, cluster_plate, n_clusters, ndims, ndims_parents):
        self.raw = distribution
        # the mixture itself has no cluster axis: a distribution that carries per-plate constants (the number of trials
        # of a multinomial) gives them up along that axis (mixture.py:37-46)
        try:
            self.squeezed = distribution.squeeze(cluster_plate) if hasattr(distribution, "squeeze") else distribution
        except ValueError as err:
            raise ValueError("Cannot mix over plate axis {0}: {1}".format(cluster_plate, str(err))) from err
        self.cluster_plate = cluster_plate
        self.K = n_clusters
        self.ndims = ndims
        self.ndims_parents = ndims_parents

    # ---- helpers --------------------------------------------------------------------------------
    def _cluster_keys(self, a, ndim):
        """Key list for an array whose plates contain the cluster axis at ``cluster_plate``."""
        npl = a.ndim - ndim
        keys = []
        for ax in range(npl):
            j = npl - ax                      # position from the right among plates (1-based)
            if -j == self.cluster_plate:
                keys.append("k")
            else:
                # plate position once the cluster axis is removed
                jj = j - 1 if j > -self.cluster_plate else j
                keys.append(("p", jj))
        return keys + [("d", i) for i in range(ndim)]

    @staticmethod
    def _plain_keys(a, ndim):
        npl = a.ndim - ndim
        return [("p", j) for j in range(npl, 0, -1)] + [("d", i) for i in range(ndim)]

    def _p_keys(self, p):
        npl = p.ndim - 1
        return [("p", j) for j in range(npl, 0, -1)] + ["k"]

    def _weighted_average(self, arr, ndim, p):
        """sum_k p[...,k] arr[...k...]  ->  plates (without cluster axis) + dims."""
        if arr.ndim - ndim < -self.cluster_plate:
            arr = arr.add_leading(-self.cluster_plate - (arr.ndim - ndim))
            # the cluster axis is then a broadcast axis: extent 1
        ak = self._cluster_keys(arr, ndim)
        pk = self._p_keys(p)
        npl = max(len([k for k in ak if isinstance(k, tuple) and k[0] == "p"]), p.ndim - 1)
        out = [("p", j) for j in range(npl, 0, -1)] + [("d", i) for i in range(ndim)]
        return D.sum_product([p, arr], [pk, ak], out)

    # ---- protocol ----------------------------------------------------------------------------------
    def compute_phi_from_parents(self, u_z, *u_params, mask=True):
        Phi = self.raw.compute_phi_from_parents(*u_params)
        p = u_z[0]
        out = []
        for Phi_i, nd in zip(Phi, self.ndims):
            Phi_i = D.asarray(Phi_i)
            if getattr(self.raw, "zero_times_inf", False):
                # mixture.py:228-230: a cluster with assignment probability exactly 0 does not count even when its
                # parameter is -inf (log of a zero probability).  Only the discrete mixed distributions can have that.
                pd = D.asarray(p)
                if Phi_i.ndim - nd < -self.cluster_plate:
                    Phi_i = Phi_i.add_leading(-self.cluster_plate - (Phi_i.ndim - nd))
                # p (plates.., K) seen with K at the cluster position and unit axes for the variable dims
                npl = pd.ndim - 1
                pos = npl + self.cluster_plate + 1                # where the cluster axis goes among the plates
                if pos < 0:
                    pd, npl, pos = pd.add_leading(-pos), npl - pos, 0
                order = list(range(pos)) + [npl] + list(range(pos, npl))
                pv = DArray(pd.owner, pd.ptr, [pd.shape[i] for i in order], [pd.strides[i] for i in order], pd.dtype)
                Phi_i = D.nonzero_select(pv.add_trailing(nd), Phi_i)
            out.append(self._weighted_average(Phi_i, nd, p))
        return out

    def compute_cgf_from_parents(self, u_z, *u_params):
        g = D.asarray(self.raw.compute_cgf_from_parents(*u_params))
        return self._weighted_average(g, 0, u_z[0])

    def compute_moments_and_cgf(self, phi, mask=True):
        return self.squeezed.compute_moments_and_cgf(phi, mask=mask)

    def compute_fixed_moments_and_f(self, x, mask=True):
        return self.squeezed.compute_fixed_moments_and_f(x, mask=True)

    def compute_message_to_parent(self, parent, index, u, u_z, *u_params):
        if index == 0:
            # L[..., k] = g_k + sum_i <phi_k,i , u_i>
            g = D.asarray(self.raw.compute_cgf_from_parents(*u_params))
            Phi = self.raw.compute_phi_from_parents(*u_params)
            if g.ndim < -self.cluster_plate:
                g = g.add_leading(-self.cluster_plate - g.ndim)
            gk = self._cluster_keys(g, 0)
            npl = max([len(gk) - 1] + [dense(ui).ndim - nd for ui, nd in zip(u, self.ndims)])
            out = [("p", j) for j in range(npl, 0, -1)] + ["k"]
            L = D.sum_product([g], [gk], out, sizes={"k": self.K})
            for Phi_i, ui, nd in zip(Phi, u, self.ndims):
                Phi_i = D.asarray(Phi_i)
                ui = dense(ui)
                if Phi_i.ndim - nd < -self.cluster_plate:
                    Phi_i = Phi_i.add_leading(-self.cluster_plate - (Phi_i.ndim - nd))
                if getattr(self.raw, "zero_times_inf", False):
                    # expfamily.py:52-58: where the moment is zero the (possibly infinite) parameter does not count —
                    # e.g. a class with probability exactly 0 that was not observed.  Only the discrete mixed
                    # distributions need it; it materialises (plates, K, dims).
                    npl_u = ui.ndim - nd
                    need = -self.cluster_plate - 1
                    ue = ui.add_leading(need - npl_u) if npl_u < need else ui
                    ue = ue.expand_dims(ue.ndim - nd + self.cluster_plate + 1)
                    sel = D.nonzero_select(ue, Phi_i)
                    t = D.sum_product([sel, ue], [self._cluster_keys(sel, nd), self._cluster_keys(ue, nd)], out,
                                      sizes={"k": self.K})
                else:
                    t = D.sum_product([Phi_i, ui], [self._cluster_keys(Phi_i, nd), self._plain_keys(ui, nd)], out,
                                      sizes={"k": self.K})
                L = D.add(L, t)
            return [L]
        # parameters of the mixed distribution
        ip = index - 1
        # u with a unit cluster axis inserted among the plates
        u_self = []
        for ui, nd in zip(u, self.ndims):
            ui = dense(ui)
            npl = ui.ndim - nd
            if npl >= -self.cluster_plate - 1:
                ui = ui.expand_dims(npl + self.cluster_plate + 1)
            u_self.append(ui)
        m = self.raw.compute_message_to_parent(parent, ip, u_self, *u_params)
        # weights: p with the cluster axis moved to its plate position
        p = u_z[0]
        need = -self.cluster_plate
        if p.ndim < need:
            p = p.add_leading(need - p.ndim)
        nd_p = p.ndim
        order = list(range(nd_p - 1))
        order.insert(nd_p + self.cluster_plate, nd_p - 1)
        pm = DArray(p.owner, p.ptr, [p.shape[i] for i in order], [p.strides[i] for i in order], p.dtype)
        w = pm
        out = []
        for mi, nd in zip(m, self.ndims_parents[ip]):
            if mi is None:
                out.append(None)
                continue
            # the mixed distribution may map plates to parent axes (e.g. GaussianARD appends its shape)
            wi = self._map_weights(ip, w)
            out.append(D.mul(D.asarray(mi), wi.add_trailing(nd)))
        return out

    def _map_weights(self, ip, w):
        extra = len(self.raw.plates_to_parent(ip, ())) if hasattr(self.raw, "plates_to_parent") else 0
        return w.add_trailing(extra) if extra else w

    def compute_weights_to_parent(self, index, weights):
        if index == 0:
            return weights
        w = np.asarray(weights)
        if w.ndim >= -self.cluster_plate:
            w = np.expand_dims(w, axis=self.cluster_plate)
        return self.raw.compute_weights_to_parent(index - 1, w)

    def plates_to_parent(self, index, plates):
        if index == 0:
            return tuple(plates)
        plates = list(plates)
        plates.insert(len(plates) + self.cluster_plate + 1, self.K)
        return tuple(self.raw.plates_to_parent(index - 1, tuple(plates)))

    def plates_from_parent(self, index, plates):
        if index == 0:
            return tuple(plates)
        plates = list(self.raw.plates_from_parent(index - 1, tuple(plates)))
        if len(plates) >= -self.cluster_plate:
            plates.pop(self.cluster_plate)
        return tuple(plates)

    def random(self, *phi, plates=None):
        return self.squeezed.random(*phi, plates=plates)

    def compute_gradient(self, g, u, phi):
        return self.squeezed.compute_gradient(g, u, phi)


class Mixture(ExponentialFamily):
    """``Mixture(z, NodeClass, *params, cluster_plate=-1, plates=None, name="")`` (mixture.py:359-488)."""

    def __init__(self, z, node_class, *params, cluster_plate=-1, plates=None, name="", initialize=True,
                 plates_multiplier=None):
        if cluster_plate >= 0:
            raise ValueError("Cluster plate axis must be negative")
        # build (and discard) a template node of the mixed class to obtain its parents / distribution
        tmpl = node_class(*params, initialize=False)
        for i, p in enumerate(tmpl.parents):
            p._remove_child(tmpl, i)
        parents = list(tmpl.parents)
        raw = tmpl._distribution
        mix_plates = list(tmpl.plates)
        if len(mix_plates) < -cluster_plate:
            raise ValueError("The mixed distribution does not have a plates axis for the cluster plate axis")
        K = mix_plates.pop(cluster_plate)
        if isinstance(z, Node) and hasattr(z, "_to_categorical"):
            z = z._to_categorical()         # e.g. a categorical Markov chain seen as categorical variables over time
        if isinstance(z, Node):
            if z.moment_kind != "categorical":
                raise ValueError("z must be a categorical-like node")
        else:
            z = categorical_constant(z, K)
        if z.dims[0][0] != K:
            raise ValueError("Inconsistent number of clusters")
        self.cluster_plate = cluster_plate
        self.moment_kind = tmpl.moment_kind
        self.mixed_class = node_class
        ndims = [len(d) for d in tmpl.dims]
        ndims_parents = [[len(d) for d in p.dims] for p in parents]
        dist = MixtureDistribution(raw, cluster_plate, K, ndims, ndims_parents)
        super().__init__(z, *parents, dims=tmpl.dims, distribution=dist, plates=plates, name=name,
                         initialize=initialize, plates_multiplier=plates_multiplier)

    def get_moments(self):
        return [dense(ui) for ui in self.u]

    def integrated_logpdf_from_parents(self, x, index):
        """log of the predictive density of ``x`` with the cluster assignment integrated out and the cluster parameters
        averaged in log scale, log sum_k p_k exp(<log p(x | theta_k)>) (mixture.py:491-545).  A query for plots and
        model checks: the result comes back as a host array with the plates of ``x``."""
        if index != 0:
            raise NotImplementedError()
        dist = self._distribution
        u, f = dist.squeezed.compute_fixed_moments_and_f(x)
        u_parents = self.moments_from_parents()
        L = dist.compute_message_to_parent(self.parents[0], 0, [dense(ui) for ui in u], *u_parents)[0]   # (.., K)
        with np.errstate(divide="ignore"):
            w = D.add(L, D.log(D.asarray(u_parents[0][0])))
        w = D.asarray(w).contiguous()
        K = w.shape[-1]
        P = tuple(w.shape[:-1])
        from .. import _bpk
        soft, g = DArray.empty(P + (K,)), DArray.empty(P)
        _bpk.get().softmax_moments(w.ptr, int(np.prod(P, dtype=np.int64)) if P else 1, K, soft.ptr, g.ptr)
        return np.asarray(D.sub(D.asarray(f), g))            # g = -logsumexp


def MultiMixture(thetas, *mixture_args, **kwargs):
    """A mixture over several cluster axes with as many categorical variables (mixture.py:548-566): nested ``Mixture``
    nodes, the i-th assignment given i trailing unit plates so that the mixed axes stay separate."""
    thetas = [t if isinstance(t, Node) else np.asarray(t) for t in thetas]
    N = len(thetas)
    thetas = [t[(Ellipsis,) + i * (None,)] for i, t in enumerate(thetas)]
    args = thetas[:1]
    for t in thetas[1:]:
        args += [Mixture, t]
    return Mixture(*(args + list(mixture_args)), **kwargs)
