"""Exponential-family stochastic nodes with device-resident natural parameters.

Role of the reference's ``nodes/stochastic.py`` (Stochastic :83-376) and
``nodes/expfamily.py`` (ExponentialFamily :94-542).  A node owns

    phi  natural parameters of q   (list of DArray, shapes broadcastable to plates+dims)
    u    moments <u(x)>_q          (list of DArray)
    g    log-normaliser of q       (DArray broadcastable to plates)
    f    base-measure term for observed plates

and a *distribution* object implementing the reference's five-method protocol
(expfamily.py:26-43) on device arrays:

    compute_phi_from_parents(*u_parents)            -> [phi_i]
    compute_cgf_from_parents(*u_parents)            -> g_p
    compute_moments_and_cgf(phi)                    -> ([u_i], g)
    compute_message_to_parent(parent, index, u, *u_parents) -> [m_i | None]
    compute_fixed_moments_and_f(x)                  -> ([u_i], f)
"""
import numpy as np

from .. import darray as D
from ..darray import DArray
from .node import Node, mask_is_full


class Distribution:
    """Protocol base (stochastic.py:16-80, expfamily.py:17-70)."""

    def compute_message_to_parent(self, parent, index, u_self, *u_parents):
        raise NotImplementedError

    def compute_phi_from_parents(self, *u_parents, mask=True):
        raise NotImplementedError

    def compute_cgf_from_parents(self, *u_parents):
        raise NotImplementedError

    def compute_moments_and_cgf(self, phi, mask=True):
        raise NotImplementedError

    def compute_fixed_moments_and_f(self, x, mask=True):
        raise NotImplementedError

    def compute_weights_to_parent(self, index, weights):
        return weights

    def compute_gradient(self, g, u, phi):
        """Euclidean gradient with respect to the natural parameters, given the Riemannian one (expfamily.py:64-70):
        the Fisher information d<u>/dphi applied to ``g``."""
        raise NotImplementedError("compute_gradient is not implemented for %s" % type(self).__name__)

    def plates_to_parent(self, index, plates):
        return plates

    def plates_from_parent(self, index, plates):
        return plates


class ExponentialFamily(Node):
    _distribution = None

    def __init__(self, *parents, dims, distribution, plates=None, name="", initialize=True, plates_multiplier=None):
        self._distribution = distribution
        self._id = Node._id_counter
        Node._id_counter += 1
        super().__init__(*parents, dims=dims, plates=plates, name=name, plates_multiplier=plates_multiplier)
        self._check_independent_parents()
        self.ndims = [len(d) for d in self.dims]
        self.observed = False            # False, True or a host bool array over plates
        self.annealing = 1.0
        self.u = [None] * len(self.dims)
        self.phi = [None] * len(self.dims)
        self.g = None
        self.f = None
        if initialize:
            self.initialize_from_prior()

    # the moments: assigning the list (also from outside, as the reference's test utilities do) invalidates every
    # cache keyed on this node's version
    @property
    def u(self):
        return self._u

    @u.setter
    def u(self, value):
        if isinstance(value, (list, tuple)):
            # host arrays handed in from outside become device arrays here, once
            value = [D.asarray(v) if isinstance(v, (np.ndarray, np.generic, float, int)) else v for v in value]
        self._u = value
        self._version = getattr(self, "_version", 0) + 1

    # ---- graph -------------------------------------------------------------------------------
    def _ids(self):
        return [self._id]

    def _plates_from_parent(self, index):
        return tuple(self._distribution.plates_from_parent(index, tuple(self.parents[index].plates)))

    def _plates_to_parent(self, index):
        return tuple(self._distribution.plates_to_parent(index, tuple(self.plates)))

    def _map_parent_axes(self, index, values):
        return tuple(self._distribution.plates_from_parent(index, tuple(values)))

    def _weights_to_parent(self, index, mask):
        return self._distribution.compute_weights_to_parent(index, mask)

    def _set_mask(self, mask):
        super()._set_mask(np.logical_or(mask, self.observed))

    # ---- moments --------------------------------------------------------------------------------
    def get_moments(self):
        return list(self.u)

    def _fully_observed(self):
        return bool(np.all(self.observed))

    def _store(self, u, g, update_mask):
        """Write new moments on the plates selected by ``update_mask`` (host bool;
        stochastic.py:223-273 np.copyto(where=mask))."""
        if mask_is_full(update_mask) or self.u[0] is None:
            self.u = [self._trim_leading(D.asarray(ui), i) for i, ui in enumerate(u)]
            if g is not None:
                self.g = g
        else:
            md = DArray.from_numpy(np.asarray(update_mask, dtype=bool), "u1")
            for i, ui in enumerate(u):
                self.u[i] = D.where(md.add_trailing(self.ndims[i]), ui, self.u[i])
            if g is not None and self.g is not None:
                self.g = D.where(md, g, self.g)
        self._version += 1

    def _trim_leading(self, a, i):
        """Moments carry exactly len(plates) + ndim_i axes: unit axes a kernel produced in front of them are dropped."""
        want = len(self.plates) + self.ndims[i]
        if isinstance(a, DArray) and a.ndim > want and all(n == 1 for n in a.shape[:a.ndim - want]):
            return a.squeeze_leading(want)
        return a

    def _canonical_phi(self, phi):
        """Give every phi_i exactly len(plates)+ndim_i axes (expfamily.py:230-250)."""
        out = []
        for i, p in enumerate(phi):
            p = D.asarray(p)
            want = len(self.plates) + self.ndims[i]
            if p.ndim < want:
                p = p.add_leading(want - p.ndim)
            elif p.ndim > want:
                p = p.squeeze_leading(want)
            out.append(p)
        return out

    def initialize_from_prior(self):
        if self._fully_observed():
            return
        u_parents = self.moments_from_parents()
        self.phi = self._canonical_phi(self._distribution.compute_phi_from_parents(*u_parents))
        u, g = self._distribution.compute_moments_and_cgf(self.phi)
        self._store(u, g, np.logical_not(self.observed))

    def initialize_from_parameters(self, *args):
        raise NotImplementedError

    def initialize_from_value(self, x, *args):
        u, _ = self._distribution.compute_fixed_moments_and_f(x, *args)
        self._store(u, None, np.logical_not(self.observed))
        self.g = D.asarray(np.inf)      # expfamily.py:206: cgf unknown until the next update

    def initialize_from_random(self):
        self.initialize_from_value(self.random())

    def random(self):
        """Draw from q on the HOST with NumPy's legacy global RNG so that seeded
        initialisations reproduce the reference bit-for-bit (SURVEY §7 'Host RNG parity')."""
        phi = [p.numpy() for p in self.phi]
        return self._distribution.random(*phi, plates=self.plates)

    # ---- the VB update (stochastic.py:276-282, expfamily.py:343-366) ---------------------------------
    def update(self, annealing=1.0):
        if self._fully_observed():
            return
        u_parents = self.moments_from_parents()
        m_children = self.message_from_children()
        phi = self._distribution.compute_phi_from_parents(*u_parents)
        phi = self._canonical_phi(phi)
        new = []
        for p, m in zip(phi, m_children):
            if m is not None:
                p = D.add(p, m) if annealing == 1.0 else D.axpby(1.0, p, annealing, m)
            if self.annealing != 1.0:
                p = D.mul(p, self.annealing)
            new.append(p)
        self.phi = self._canonical_phi(new)
        u, g = self._distribution.compute_moments_and_cgf(self.phi)
        self._store(u, g, np.logical_not(self.observed))

    # ---- parameters and gradients (expfamily.py:260-340): the natural parameters are the optimisation variables ------
    def get_parameters(self):
        from .gaussian import dense
        return [D.asarray(dense(p)).contiguous().copy() for p in self.phi]

    def set_parameters(self, x):
        self._fused = None            # a fused sweep's shortcuts (shared covariance, virtual phi0 / g) no longer describe q
        self.phi = self._canonical_phi([D.asarray(p) for p in x])
        u, g = self._distribution.compute_moments_and_cgf(self.phi)
        self._store(u, g, np.logical_not(self.observed))

    def get_riemannian_gradient(self):
        """Natural gradient of the bound: annealing * (<phi>_prior + sum of the children's messages) - phi."""
        from .gaussian import dense
        u_parents = self.moments_from_parents()
        m_children = self.message_from_children()
        phi_p = self._canonical_phi(self._distribution.compute_phi_from_parents(*u_parents))
        out = []
        for i, (p, m) in enumerate(zip(phi_p, m_children)):
            t = D.add(p, m) if m is not None else p
            t = D.axpby(self.annealing, t, -1.0, D.asarray(dense(self.phi[i])))
            out.append(t.broadcast_to(self.get_shape(i)).contiguous())
        return out

    def get_gradient(self, rg):
        """Gradient with respect to the natural parameters; takes the Riemannian gradient as input
        (expfamily.py:283-297)."""
        from .gaussian import dense
        g = self._distribution.compute_gradient(rg, [D.asarray(dense(ui)) for ui in self.u],
                                                [D.asarray(dense(p)) for p in self.phi])
        if self.annealing != 1.0:
            g = [D.mul(gi, 1.0 / self.annealing) for gi in g]
        return g

    def logpdf(self, X, mask=True):
        """log q(X) of this node's current distribution (expfamily.py:483-497); a host array over the plates."""
        if mask is not True:
            raise NotImplementedError("Mask not yet implemented")
        from .gaussian import dense
        u, f = self._distribution.compute_fixed_moments_and_f(X, mask=mask)
        Z = D.add(D.asarray(self.g.materialize() if hasattr(self.g, "materialize") else self.g), D.asarray(f))
        for phi_d, u_d, nd in zip(self.phi, u, self.ndims):
            phi_d = D.asarray(dense(phi_d))
            u_d = D.asarray(dense(u_d))
            npl = max(phi_d.ndim, u_d.ndim) - nd
            keys_p = list(range(npl))
            keys_d = list(range(100, 100 + nd))
            Z = D.add(Z, D.sum_product([phi_d, u_d], [keys_p[npl - (phi_d.ndim - nd):] + keys_d,
                                                       keys_p[npl - (u_d.ndim - nd):] + keys_d], keys_p))
        return np.asarray(Z)

    def pdf(self, X, mask=True):
        return np.exp(self.logpdf(X, mask=mask))

    def observe(self, x, *args, mask=True):
        """Fix the moments on the observed plates and propagate the mask
        (expfamily.py:369-398)."""
        u, f = self._distribution.compute_fixed_moments_and_f(x, *args, mask=mask)
        for i, ui in enumerate(u):
            if tuple(ui.shape) != self.get_shape(i):
                raise ValueError("Shape of the given array not equal to the shape of the node.\n"
                                 "Received shape: %s\nExpected shape: %s\nCheck plates."
                                 % (tuple(ui.shape), self.get_shape(i)))
        mask = np.asarray(mask, dtype=bool) if mask is not True else True
        if mask is not True and mask.ndim > len(self.plates):
            raise ValueError("mask has more axes than the node has plates")
        if mask_is_full(mask):
            self.u = list(u)
            self.f = f
            self.observed = True
        else:
            md = DArray.from_numpy(mask, "u1")
            for i, ui in enumerate(u):
                old = self.u[i] if self.u[i] is not None else D.asarray(np.nan)
                self.u[i] = D.where(md.add_trailing(self.ndims[i]), ui, old)
            self.f = f
            self.observed = mask
        self._version += 1
        self._update_mask()

    def unobserve(self):
        self.observed = False
        self._update_mask()

    def _message_and_mask_to_parent(self, index):
        u_parents = self.moments_from_parents(exclude=index)
        m = self._distribution.compute_message_to_parent(self.parents[index], index, self.get_moments(),
                                                         *u_parents)
        mask = self._distribution.compute_weights_to_parent(index, self.mask)
        return m, mask

    # ---- lower bound (expfamily.py:400-480) ----------------------------------------------------------
    def lower_bound_contribution(self, ignore_masked=True):
        """E[log p(X|parents)] - E[log q(X)] summed over the active plates -> 0-d DArray.  ``ignore_masked=False`` sums
        over ALL plates, also those no observation depends on (expfamily.py:470-480)."""
        u_parents = self.moments_from_parents()
        phi_p = self._canonical_phi(self._distribution.compute_phi_from_parents(*u_parents))
        L = D.asarray(self._distribution.compute_cgf_from_parents(*u_parents))
        all_obs = self._fully_observed()
        none_obs = not np.any(self.observed)
        T = 1.0 / self.annealing
        if all_obs:
            L = D.add(L, self.f)
        elif none_obs:
            L = D.sub(L, self.g if T == 1.0 else D.mul(self.g, T))
        else:
            obs = DArray.from_numpy(np.asarray(self.observed, dtype=bool), "u1")
            z = D.affine(self.g, -T)
            L = D.add(L, D.where(obs, D.asarray(self.f), z))
        for i in range(len(self.dims)):
            nd = self.ndims[i]
            if all_obs:
                diff = phi_p[i]
            elif none_obs:
                diff = D.axpby(1.0, phi_p[i], -T, self.phi[i])
            else:
                lat = DArray.from_numpy(np.logical_not(self.observed), "u1").add_trailing(nd)
                diff = D.axpby(1.0, phi_p[i], -T, D.where(lat, self.phi[i], D.asarray(0.0)))
            ui = D.asarray(self.u[i])
            if self._guard_zero_times_inf:
                diff = D.nonzero_select(ui, diff)
            nplate = max(ui.ndim, diff.ndim) - nd
            keys_p = list(range(nplate))
            keys_d = list(range(100, 100 + nd))
            a_keys = keys_p[nplate - (diff.ndim - nd):] + keys_d
            b_keys = keys_p[nplate - (ui.ndim - nd):] + keys_d
            Z = D.sum_product([diff, ui], [a_keys, b_keys], keys_p)
            L = D.add(L, Z)
        mdev = self.mask_device() if ignore_masked else None
        scale = float(np.prod(self.plates_multiplier)) if self.plates_multiplier else 1.0
        return D.reduce_to_shape(L, (), mask=mdev, from_shape=self.plates, scale=scale)

    _guard_zero_times_inf = False

    # ---- convenience -----------------------------------------------------------------------------------
    def get_parameters(self):
        return [p for p in self.phi]

    def __str__(self):
        return "%s(%s)" % (type(self).__name__, self.name)
