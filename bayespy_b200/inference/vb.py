"""VB — the variational Bayes sweep scheduler (replaces ``VB.update``,
``_end_iteration_step`` and ``loglikelihood_lowerbound`` of
bayespy/inference/vmp/vmp.py:52-232, :682-764).

Host side: node order, ``repeat`` / ``tol`` / ``verbose`` semantics, the
iteration print format (vmp.py:725, doctests match on it), the lower-bound
histories ``L``, ``l`` and ``cputime``.  Device side: every node update and
every lower-bound term is enqueued on one CUDA stream; the only device->host
traffic per sweep is the vector of per-node bound terms (a few doubles), which
the convergence test needs (vmp.py:717-747).

When the model contains a sub-graph a fused sweep kernel exists for (the
plated Gaussian factor model of PCA, the Gaussian mixture), ``VB`` attaches the
corresponding *plan* (``bayespy_b200.engine.plans``), which serves the same node
updates from one pass over the data instead of one pass per message.
"""
import time
import warnings

import numpy as np

from .. import darray as D
from ..darray import DArray
from ..engine.node import Node


class VB:

    def __init__(self, *nodes, tol=1e-5, autosave_filename=None, autosave_iterations=0,
                 use_logging=False, user_data=None, callback=None, fused=True, resident=True):
        self.user_data = user_data
        for ind, node in enumerate(nodes):
            if not isinstance(node, Node):
                raise ValueError("Argument number %d is not a node" % (ind + 1))
        if use_logging:
            import logging
            self.print = logging.getLogger(__name__).info
        else:
            self.print = print
        model = []
        for n in nodes:
            if n not in model:
                model.append(n)
        self.model = model
        self.ignore_bound_checks = False
        self.iter = 0
        self.annealing_changed = False
        self.converged = False
        self.L = np.array(())
        self.cputime = np.array(())
        self.l = {node: np.array([]) for node in self.model}
        self.autosave_iterations = autosave_iterations
        if not autosave_filename:
            # like the reference (vmp.py:86-97): a dated file name in the temporary directory, so that save() / load()
            # without a name work; no file is created before the first save
            import datetime
            import tempfile
            date = datetime.datetime.today().strftime("%Y%m%d%H%M%S")
            fd = tempfile.NamedTemporaryFile(prefix="vb_autosave_%s_" % date, suffix=".hdf5", delete=True)
            autosave_filename = fd.name
            fd.close()
        self.autosave_filename = autosave_filename
        names = [node.name for node in self.model]
        if len(names) != len(self.model):
            raise Exception("Use unique names for nodes.")
        self.callback = callback
        self.callback_output = None
        self.tol = tol
        self.resident = resident        # allow whole sweeps to stay on the device (plans.run_resident)
        self.plans = []
        if fused:
            from ..engine import plans
            self.plans = plans.attach(self.model)

    # ---- container protocol (vmp.py:358-400) ---------------------------------------------------
    def __getitem__(self, name):
        if isinstance(name, Node):
            return name
        for node in self.model:
            if node.name == name:
                return node
        raise ValueError("Node %s not found" % (name,))

    def use_logging(self, use):
        """Send the iteration lines to the ``logging`` module instead of printing them (vmp.py:111-118)."""
        if use:
            import logging
            self.print = logging.getLogger(__name__).info
        else:
            self.print = print

    def set_autosave(self, filename, iterations=None, nodes=None):
        """vmp.py:121-126."""
        self.autosave_filename = filename
        self.filename = filename
        self.autosave_nodes = nodes
        if iterations is not None:
            self.autosave_iterations = iterations

    def plot(self, *nodes, **kwargs):
        """Plot the given nodes (default: every node that has a plotter), vmp.py:767-790."""
        nodes = self.model if len(nodes) == 0 else [self[n] for n in nodes if n is not None]
        for node in nodes:
            if node.has_plotter():
                node.plot(**kwargs)

    # ---- checkpoints (vmp.py:237-356; container and hierarchy: inference/checkpoint.py) ------------------------
    def save(self, *nodes, filename=None):
        from . import checkpoint
        nodes = self.model if len(nodes) == 0 else [self[n] for n in nodes if n is not None]
        if not filename:
            if self.autosave_filename:
                filename = self.autosave_filename
            else:
                raise Exception("Filename must be given.")
        w = checkpoint.open_writer(filename)
        try:
            for node in nodes:
                if node.name == "":
                    raise Exception("In order to save nodes, they must have (unique) names.")
                if hasattr(node, "u") and hasattr(node, "observed"):
                    checkpoint.save_node(w, node, "nodes/%s" % node.name)
            w.put("L", self.L)
            w.put("cputime", self.cputime)
            w.put("iter", self.iter)
            w.put("converged", self.converged)
            if self.callback_output is not None:
                w.put("callback_output", self.callback_output)
            for node in nodes:
                w.put("boundterms/%s" % node.name, self.l[node])
            if self.user_data is not None:
                for key, value in self.user_data.items():
                    w.put("user_data/%s" % key, value)
        finally:
            w.close()

    def load(self, *nodes, filename=None, nodes_only=False):
        from . import checkpoint
        if not filename:
            if self.autosave_filename:
                filename = self.autosave_filename
            else:
                raise Exception("Filename must be given.")
        r = checkpoint._Reader(filename)
        try:
            nodes = self.model if len(nodes) == 0 else [self[n] for n in nodes if n is not None]
            for node in nodes:
                if node.name == "":
                    raise Exception("In order to load nodes, they must have (unique) names.")
                if hasattr(node, "u") and hasattr(node, "observed"):
                    checkpoint.load_node(r, node, "nodes/%s" % node.name)
            for plan in self.plans:                      # cached statistics / device state describe the old posterior
                for attr in ("_stats", "_mstats", "_e2", "_me2", "_res_cache"):
                    if hasattr(plan, attr):
                        setattr(plan, attr, None)
            if not nodes_only:
                self.L = r.get("L")
                self.cputime = r.get("cputime")
                self.iter = int(r.get("iter"))
                self.converged = bool(r.get("converged"))
                for node in nodes:
                    self.l[node] = r.get("boundterms/%s" % node.name)
                if r.has("callback_output"):
                    self.callback_output = r.get("callback_output")
        finally:
            r.close()

    def set_callback(self, callback):
        self.callback = callback

    def has_converged(self, tol=None):
        return self.converged

    # ---- the sweep (vmp.py:132-172) ------------------------------------------------------------
    def update(self, *nodes, repeat=1, plot=False, tol=None, verbose=True, tqdm=None):
        if len(nodes) == 0:
            nodes = self.model
        # whole sweeps that can stay on the device do (engine/plans.py: FactorModelPlan.run_resident)
        sweep_plan = None
        if self.resident and self.autosave_iterations == 0:
            for plan in self.plans:
                prog = plan.resident_program(self, nodes) if hasattr(plan, "resident_program") else None
                if prog is None:
                    continue
                if tqdm is None and not plot and not callable(self.callback):
                    plan.run_resident(self, prog, repeat, tol, verbose)
                    return
                # a callback / progress bar / plot wants the host between the node updates and the bound
                # (vmp.py:702-713): keep every node update of the sweep in ONE fused launch, then hand over
                if hasattr(plan, "sweep_resident"):
                    sweep_plan = (plan, prog)
                break
        if tqdm is not None:
            tqdm = tqdm(total=repeat)
        i = 0
        while repeat is None or i < repeat:
            t = time.time()
            if sweep_plan is not None:
                # the program is re-derived every iteration: the callback may have changed the graph
                prog = sweep_plan[0].resident_program(self, nodes)
                if prog is None:
                    sweep_plan = None
            if sweep_plan is not None:
                sweep_plan[0].sweep_resident(self, prog)
            else:
                for node in nodes:
                    X = self[node]
                    if hasattr(X, "update") and callable(X.update):
                        X.update()
            cputime = time.time() - t
            i += 1
            if tqdm is not None:
                tqdm.update()
            if self._end_iteration_step(None, cputime, tol=tol, verbose=verbose):
                return

    # ---- lower bound (vmp.py:180-199) -----------------------------------------------------------
    def _bound_terms(self, ignore_masked=True):
        """One D2H of len(model) doubles: every node's term is produced on device."""
        if ignore_masked:
            terms = [node.lower_bound_contribution() for node in self.model]
        else:
            terms = [node.lower_bound_contribution(ignore_masked=False) for node in self.model]
        vec = DArray.empty((len(terms),))
        host_fill = {}
        for i, t in enumerate(terms):
            if isinstance(t, DArray):
                D.copy_into(vec.slice_axis(0, i, i + 1), t.reshape((1,)))
            else:
                host_fill[i] = float(t)
        out = vec.numpy()
        for i, v in host_fill.items():
            out[i] = v
        return out

    def compute_lowerbound(self, ignore_masked=True):
        return float(np.sum(self._bound_terms(ignore_masked)))

    def compute_lowerbound_terms(self, *nodes):
        vals = self._bound_terms()
        d = dict(zip(self.model, vals))
        if len(nodes) == 0:
            return d
        return {self[n]: d[self[n]] for n in nodes}

    def loglikelihood_lowerbound(self):
        vals = self._bound_terms()
        for node, lp in zip(self.model, vals):
            self.l[node][self.iter] = lp
        return float(np.sum(vals))

    def get_iteration_by_nodes(self):
        return self.l

    def _append_iterations(self, iters):
        nans = np.full(iters, np.nan)
        self.L = np.append(self.L, nans)
        self.cputime = np.append(self.cputime, nans)
        for node in self.l:
            self.l[node] = np.append(self.l[node], nans)

    def _record_resident_iterations(self, rows, terms_of, dt, stop, check, verbose):
        """Book-keeping of ``n`` iterations a device-resident loop has run (one row of bound terms each,
        total in column 5): what _end_iteration_step does per iteration (vmp.py:693-764), for a whole chunk with
        array assignments — the per-iteration Python loop is kept for verbose runs only."""
        import warnings
        n = len(rows)
        if n == 0:
            return
        while self.iter + n > len(self.L):
            self._append_iterations(100)
        i0 = self.iter
        Ls = rows[:, 5]
        for node in self.model:
            self.l[node][i0:i0 + n] = rows[:, terms_of[node]] if node in terms_of else 0.0
        self.L[i0:i0 + n] = Ls
        self.cputime[i0:i0 + n] = dt / n
        compare = check and not self.annealing_changed
        if verbose:
            for r in range(n):
                self.print("Iteration %d: loglike=%e (%.3f seconds)" % (i0 + r + 1, Ls[r], dt / n))
        self.converged = False
        if compare:
            prev = np.concatenate(([self.L[i0 - 1]] if i0 > 0 else [np.nan], Ls[:-1]))
            if i0 == 0:
                prev[0] = np.nan
            for d in (prev - Ls)[(prev - Ls) > 1e-6]:
                warnings.warn("Lower bound decreased %e! Bug somewhere or numerical inaccuracy?" % d)
            if stop and (i0 + n - 1) > 0:
                if verbose:
                    self.print("Converged at iteration %d." % (i0 + n))
                self.converged = True
        elif check and n > 1:
            # the first iteration after an annealing change is not compared; the later ones of the chunk are
            d = Ls[:-1] - Ls[1:]
            for v in d[d > 1e-6]:
                warnings.warn("Lower bound decreased %e! Bug somewhere or numerical inaccuracy?" % v)
            if stop:
                if verbose:
                    self.print("Converged at iteration %d." % (i0 + n))
                self.converged = True
        self.annealing_changed = False
        self.iter = i0 + n

    def _end_iteration_step(self, method, cputime, tol=None, verbose=True, bound_cpu_time=True):
        """vmp.py:693-764: callback, bound, print, decrease warning, convergence test."""
        if self.iter >= len(self.L):
            self._append_iterations(100)
        if callable(self.callback):
            z = self.callback()
            if z is not None:
                z = np.array(z)[..., np.newaxis]
                if self.callback_output is None:
                    self.callback_output = z
                else:
                    self.callback_output = np.concatenate((self.callback_output, z), axis=-1)
        t = time.time()
        L = self.loglikelihood_lowerbound()
        if bound_cpu_time:
            cputime += time.time() - t
        self.cputime[self.iter] = cputime
        self.L[self.iter] = L
        if verbose:
            if method:
                self.print("Iteration %d (%s): loglike=%e (%.3f seconds)" % (self.iter + 1, method, L, cputime))
            else:
                self.print("Iteration %d: loglike=%e (%.3f seconds)" % (self.iter + 1, L, cputime))
        self.converged = False
        if not self.ignore_bound_checks and not self.annealing_changed and self.iter > 0:
            if self.L[self.iter - 1] - L > 1e-6:
                L_diff = (self.L[self.iter - 1] - L)
                warnings.warn("Lower bound decreased %e! Bug somewhere or numerical inaccuracy?" % L_diff)
            L0 = self.L[self.iter - 1]
            L1 = self.L[self.iter]
            if tol is None:
                tol = self.tol
            div = 0.5 * (abs(L0) + abs(L1))
            if (L1 - L0) / div < tol:
                if verbose:
                    self.print("Converged at iteration %d." % (self.iter + 1))
                self.converged = True
        self.annealing_changed = False
        self.iter += 1
        # vmp.py:750-758
        if self.autosave_iterations > 0 and np.mod(self.iter, self.autosave_iterations) == 0:
            if self.autosave_filename is not None:
                self.save(filename=self.autosave_filename)
                if verbose:
                    self.print("Auto-saved to %s" % self.autosave_filename)
        return self.converged

    # ---- gradient-based learning (vmp.py:402-662): the nodes' natural parameters are the variables ----------------
    def get_gradients(self, *nodes, euclidian=False):
        """Riemannian gradients of the given nodes (and, on request, the Euclidean ones too)."""
        rg = [self[node].get_riemannian_gradient() for node in nodes]
        if euclidian:
            g = [self[node].get_gradient(rg_x) for node, rg_x in zip(nodes, rg)]
            return rg, g
        return rg

    def get_parameters(self, *nodes):
        return [self[node].get_parameters() for node in nodes]

    def set_parameters(self, x, *nodes):
        for node, xi in zip(nodes, x):
            self[node].set_parameters(xi)

    def gradient_step(self, *nodes, scale=1.0):
        """One step of natural-gradient ascent."""
        p = self.add(self.get_parameters(*nodes), self.get_gradients(*nodes), scale=scale)
        self.set_parameters(p, *nodes)

    def dot(self, x1, x2):
        """Inner product of two vectors in parameter format (one list of arrays per node)."""
        v = 0.0
        for y1, y2 in zip(x1, x2):
            for z1, z2 in zip(y1, y2):
                a, b = D.asarray(z1), D.asarray(z2)
                keys = list(range(a.ndim))
                v += float(D.sum_product([a, b.broadcast_to(a.shape)], [keys, keys], []).numpy())
        return v

    def add(self, x1, x2, scale=1):
        return [[D.axpby(1.0, D.asarray(z1), float(scale), D.asarray(z2)) for z1, z2 in zip(y1, y2)]
                for y1, y2 in zip(x1, x2)]

    def optimize(self, *nodes, maxiter=10, verbose=True, method="fletcher-reeves", riemannian=True, collapsed=None,
                 tol=None):
        """Riemannian conjugate gradient over the natural parameters of ``nodes``; the ``collapsed`` nodes are
        re-optimised in closed form after every trial step (vmp.py:470-605)."""
        method = method.lower()
        if method not in ("gradient", "fletcher-reeves"):
            raise Exception("Unknown optimization method: %s" % (method))
        collapsed = [] if collapsed is None else list(collapsed)
        scale = 1.0
        p = self.get_parameters(*nodes)
        dd_prev = 0
        s = None
        for _ in range(maxiter):
            t = time.time()
            if riemannian and method == "gradient":
                rg = self.get_gradients(*nodes, euclidian=False)
                g1 = g2 = rg
            else:
                rg, g = self.get_gradients(*nodes, euclidian=True)
                g1, g2 = (g, rg) if riemannian else (g, g)
            if method == "gradient":
                b = 0
            else:
                dd_curr = self.dot(g1, g2)
                b = 0 if dd_prev == 0 else dd_curr / dd_prev
                dd_prev = dd_curr
            s = self.add(g2, s, scale=b) if b else g2
            success = False
            while not success:
                p_new = self.add(p, s, scale=scale)
                try:
                    self.set_parameters(p_new, *nodes)
                except Exception:
                    if verbose:
                        self.print("CG update was unsuccessful, using gradient and resetting CG")
                    if s is g2:
                        scale = scale / 2
                    dd_prev = 0
                    s = g2
                    continue
                collapsed_params = self.get_parameters(*collapsed)
                try:
                    for node in collapsed:
                        self[node].update()
                except Exception:
                    self.set_parameters(collapsed_params, *collapsed)
                    if verbose:
                        self.print("Collapsed node update node failed, reset CG")
                    if s is g2:
                        scale = scale / 2
                    dd_prev = 0
                    s = g2
                    continue
                L = self.compute_lowerbound()
                bound_decreased = (self.iter > 0 and L < self.L[self.iter - 1]
                                   and not np.allclose(L, self.L[self.iter - 1], rtol=1e-8))
                if np.isnan(L) or bound_decreased:
                    self.set_parameters(collapsed_params, *collapsed)
                    if s is g2:
                        scale = scale / 2
                        if verbose:
                            self.print("Gradient ascent decreased lower bound from {0} to {1}, halfing step length"
                                       .format(self.L[self.iter - 1], L))
                    elif scale < 2 ** (-10):
                        if verbose:
                            self.print("CG decreased lower bound from {0} to {1}, reset CG."
                                       .format(self.L[self.iter - 1], L))
                        dd_prev = 0
                        s = g2
                    else:
                        scale = scale / 2
                        if verbose:
                            self.print("CG decreased lower bound from {0} to {1}, halfing step length"
                                       .format(self.L[self.iter - 1], L))
                    continue
                success = True
            scale = scale * np.sqrt(2)
            p = p_new
            cputime = time.time() - t
            if self._end_iteration_step("OPT", cputime, tol=tol, verbose=verbose):
                break

    def pattern_search(self, *nodes, collapsed=None, maxiter=3):
        """Pattern search along the direction of one VB update of ``nodes`` (vmp.py:608-662)."""
        import scipy.optimize
        collapsed = [] if collapsed is None else list(collapsed)
        t = time.time()
        for x in nodes:
            self[x].update()
        for x in collapsed:
            self[x].update()
        p0 = self.get_parameters(*nodes)
        for x in nodes:
            self[x].update()
        p1 = self.get_parameters(*nodes)
        dp = self.add(p1, p0, scale=-1)

        def cost(alpha):
            p_new = self.add(p1, dp, scale=alpha)
            try:
                self.set_parameters(p_new, *nodes)
            except Exception:
                return np.inf
            for x in collapsed:
                self[x].update()
            return -self.compute_lowerbound()

        res = scipy.optimize.minimize_scalar(cost, bracket=[0, 3], options={"maxiter": maxiter})
        self.set_parameters(self.add(p1, dp, scale=res.x), *nodes)
        for x in collapsed:
            self[x].update()
        cputime = time.time() - t
        self._end_iteration_step("PS", cputime)

    def set_annealing(self, annealing):
        """vmp.py:665-679."""
        for node in self.model:
            node.annealing = annealing
        self.annealing_changed = True
        self.converged = False
