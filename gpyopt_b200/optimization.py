"""Callers of the hot path, re-hosted so that they feed it batches instead of single points (SURVEY.md section 8f,
rows f2 / f4).  Needs the reference package (GPyOpt) importable: these classes reuse its optimiser and space code.

* :class:`BatchedAcquisitionOptimizer` -- drop-in for ``GPyOpt.optimization.AcquisitionOptimizer``
  (GPyOpt/optimization/acquisition_optimizer.py:16-78).  The reference optimises its 5 anchor points one after the
  other, and ``OptLbfgs`` evaluates ``f(x)`` and ``f_df(x)`` at a single point per step (optimizer.py:44-51): a
  latency-bound chain of N=1 device calls.  Two ways out:
  ``local_optimizer='scipy'`` (default): every anchor runs the UNCHANGED reference ``apply_optimizer`` (and therefore
  SciPy's own L-BFGS-B) in its own thread, and a rendezvous stacks the points all threads are currently asking for into
  ONE device call per round.  Each L-BFGS instance sees the values it would have seen alone: a row's result does not
  depend on the other rows of its call (bitwise for calls of up to 32 rows, ``test_small_batch_latency_path``; stacked
  calls with more rows take the tensor-core path and agree to rounding, ~1e-13), so the optimum is the reference's.
  ``local_optimizer='device'``: the optimiser itself runs inside the library (``gpo_lbfgs``): all anchors advance in
  lock-step on the GPU -- projected L-BFGS, one fused value + gradient evaluation per round, no host round trip per step;
  same stopping rules as SciPy's L-BFGS-B, optimum equal to the stopping tolerance rather than bit for bit.
* :class:`FusedAnchorPointsGenerator` -- ``ObjectiveAnchorPointsGenerator`` whose scoring + ``argsort[:k]``
  (anchor_points_generator.py:62-67) is the fused device sweep + top-k (``gpo_acq_topk``) when the acquisition has
  constant cost and the space no constraints.
* :class:`FusedThompsonAnchorPointsGenerator` -- Thompson-sampling anchors (anchor_points_generator.py:72-88) scored
  with one sweep and one vectorised normal draw instead of a Python loop over 25 000 candidates (same random stream).
* :class:`DeviceDesignAnchorPointsGenerator` -- the same, with the random design itself generated on the device
  (``gpo_acq_topk_uniform``): for all-continuous box domains the 1000 ... 1e7 candidate rows of
  ``initial_design('random', ...)`` (experiment_design/random_design.py:56-76) never exist on the host.
"""
import threading

import numpy as np

from ._compat import HAVE_GPYOPT, constant_cost_withGradients

if HAVE_GPYOPT:  # pragma: no branch
    from GPyOpt.optimization.acquisition_optimizer import (AcquisitionOptimizer, max_objective_anchor_points_logic,
                                                           thompson_sampling_anchor_points_logic, sobol_design_type,
                                                           random_design_type)
    from GPyOpt.optimization.anchor_points_generator import (ObjectiveAnchorPointsGenerator,
                                                             ThompsonSamplingAnchorPointsGenerator)
    from GPyOpt.optimization.optimizer import apply_optimizer, choose_optimizer
    from GPyOpt.experiment_design import initial_design
    from GPyOpt.core.errors import FullyExploredOptimizationDomainError
else:  # the classes below are only usable next to the reference package
    AcquisitionOptimizer = object
    ObjectiveAnchorPointsGenerator = object


class Rendezvous(object):
    """Batches the calls that several worker threads make to row-wise functions.

    ``fns`` maps a name to a batched function ``X[N,d] -> array or tuple of arrays with N leading rows``.  A worker
    calls ``call(name, x)`` with one row and blocks; when every live worker is blocked the coordinator evaluates each
    requested function ONCE on the stacked rows and hands every worker its own row of the result."""

    def __init__(self, fns, n_workers):
        self.fns = fns
        self.cv = threading.Condition()
        self.live = n_workers
        self.pending = []      # (name, x_row, slot)
        self.calls = 0         # batched evaluations issued (for tests / statistics)
        self.rows = 0

    def call(self, name, x):
        slot = {'done': False, 'out': None, 'err': None}
        with self.cv:
            self.pending.append((name, np.atleast_2d(np.asarray(x, dtype=np.float64)), slot))
            self.cv.notify_all()
            while not slot['done']:
                self.cv.wait()
        if slot['err'] is not None:
            raise slot['err']
        return slot['out']

    def worker_done(self):
        with self.cv:
            self.live -= 1
            self.cv.notify_all()

    def serve(self):
        """Run on the coordinating thread until every worker has finished."""
        while True:
            with self.cv:
                while self.live > 0 and len(self.pending) < self.live:
                    self.cv.wait()
                if self.live == 0 and not self.pending:
                    return
                batch, self.pending = self.pending, []
            by_name = {}
            for name, x, slot in batch:
                by_name.setdefault(name, []).append((x, slot))
            for name, items in by_name.items():
                rows = [x.shape[0] for x, _ in items]
                try:
                    res = self.fns[name](np.vstack([x for x, _ in items]))
                    self.calls += 1
                    self.rows += sum(rows)
                    off = 0
                    for (x, slot), r in zip(items, rows):
                        slot['out'] = (tuple(np.asarray(a)[off:off + r] for a in res) if isinstance(res, tuple)
                                       else np.asarray(res)[off:off + r])
                        off += r
                except Exception as e:  # hand the failure to the workers that asked
                    for _, slot in items:
                        slot['err'] = e
            with self.cv:
                for _, _, slot in batch:
                    slot['done'] = True
                self.cv.notify_all()


def _plain(acquisition):
    """True when acquisition_function == -_compute_acq (constant cost, no constraints, no one-hot zipping)."""
    space = acquisition.space
    return (getattr(acquisition, 'cost_withGradients', None) is constant_cost_withGradients and space is not None
            and not getattr(space, 'constraints', None) and space.model_dimensionality == space.objective_dimensionality)


class FusedAnchorPointsGenerator(ObjectiveAnchorPointsGenerator):
    """ObjectiveAnchorPointsGenerator (anchor_points_generator.py:91-104) with the scoring sweep and the top-k
    selection fused on the device.  Falls back to the reference logic whenever duplicates, context variables or a
    non-plain acquisition are involved."""

    def __init__(self, space, design_type, acquisition, num_samples=1000):
        super(FusedAnchorPointsGenerator, self).__init__(space, design_type, acquisition.acquisition_function, num_samples)
        self.acquisition = acquisition

    def get(self, num_anchor=5, duplicate_manager=None, unique=False, context_manager=None):
        from .acquisitions import _FusedAcquisition, _engine_of
        no_context = context_manager is None or len(getattr(context_manager, 'context_index', [])) == 0
        fused = (isinstance(self.acquisition, _FusedAcquisition) and _plain(self.acquisition) and not duplicate_manager
                 and not unique and no_context)
        if not fused:
            return super(FusedAnchorPointsGenerator, self).get(num_anchor, duplicate_manager, unique, context_manager)
        X = initial_design(self.design_type, self.space, self.num_samples)   # same RNG stream as the reference
        X = self.space.unzip_inputs(X)
        if X.shape[0] == 0:
            raise FullyExploredOptimizationDomainError("No anchor points could be generated")
        eng = _engine_of(self.acquisition.model)
        _, idx = eng.acq_topk(self.acquisition._kind, self.acquisition._par(), np.ascontiguousarray(X, dtype=np.float64),
                              min(num_anchor, X.shape[0]))
        return X[idx, :]


class DeviceDesignAnchorPointsGenerator(FusedAnchorPointsGenerator):
    """Anchor points from a uniform design generated, scored and reduced to its top-k on the device in one call.

    The design is the counter-based stream ``numpy.random.Generator(numpy.random.Philox(key=[seed, call])).uniform``
    (same distribution as the reference's ``samples_multidimensional_uniform``, a different stream than its legacy
    global MT19937 one); ``call`` counts the ``get`` calls so that every BO iteration sees a fresh design.  Falls back
    to the host design whenever the space is not a plain continuous box or the acquisition is not a fused one."""

    def __init__(self, space, acquisition, num_samples=1000, seed=0):
        super(DeviceDesignAnchorPointsGenerator, self).__init__(space, random_design_type, acquisition, num_samples)
        self.seed, self.calls = int(seed), 0

    def _continuous_box(self):
        types = getattr(self.space, 'has_types', {})
        return (self.space.has_continuous() and not any(v for t, v in types.items() if t != 'continuous')
                and not self.space.has_constraints())

    def get(self, num_anchor=5, duplicate_manager=None, unique=False, context_manager=None):
        from .acquisitions import _FusedAcquisition, _engine_of
        no_context = context_manager is None or len(getattr(context_manager, 'context_index', [])) == 0
        fused = (isinstance(self.acquisition, _FusedAcquisition) and _plain(self.acquisition) and not duplicate_manager
                 and not unique and no_context and self._continuous_box())
        if not fused:
            return super(DeviceDesignAnchorPointsGenerator, self).get(num_anchor, duplicate_manager, unique, context_manager)
        eng = _engine_of(self.acquisition.model)
        key = [self.seed, self.calls]
        self.calls += 1
        bounds = np.asarray(self.space.get_bounds(), dtype=np.float64)
        _, _, rows = eng.acq_topk_uniform(self.acquisition._kind, self.acquisition._par(), key, self.num_samples, bounds,
                                          min(num_anchor, self.num_samples))
        return rows


class FusedThompsonAnchorPointsGenerator(ThompsonSamplingAnchorPointsGenerator if HAVE_GPYOPT else object):
    """ThompsonSamplingAnchorPointsGenerator (anchor_points_generator.py:72-88) without its 25 000-iteration Python loop:
    one device value sweep (``model.predict``) and ONE vectorised ``numpy.random.normal`` call, which consumes NumPy's
    global legacy stream exactly like the reference's element-by-element loop -- the scores, and therefore the anchors,
    are identical for a given seed."""

    def get_anchor_point_scores(self, X):
        means, stds = self.model.predict(X)
        return np.random.normal(np.asarray(means).reshape(-1), np.asarray(stds).reshape(-1))


class BatchedAcquisitionOptimizer(AcquisitionOptimizer):
    """AcquisitionOptimizer whose local optimisations share their device calls (see module docstring)."""

    def __init__(self, space, optimizer='lbfgs', acquisition=None, device_design_seed=None, num_anchor_samples=1000,
                 local_optimizer='scipy', lbfgs_options=None, **kwargs):
        if not HAVE_GPYOPT:
            raise ImportError("BatchedAcquisitionOptimizer builds on GPyOpt's optimizer classes; GPyOpt is not importable")
        super(BatchedAcquisitionOptimizer, self).__init__(space, optimizer, **kwargs)
        self.acquisition = acquisition     # optional: enables the fused anchor sweep
        self.device_design_seed = device_design_seed   # not None: the anchor design is generated on the device
        self.num_anchor_samples = int(num_anchor_samples)
        # 'scipy': the reference's own SciPy L-BFGS-B per anchor, device calls batched by the rendezvous (bit-identical
        # optimum).  'device': all anchors advance in lock-step inside the library (gpo_lbfgs: projected L-BFGS on the GPU,
        # one fused evaluation per round, no host round trip per step) -- same stopping rules, optimum equal to tolerance.
        self.local_optimizer = local_optimizer
        # tighter than SciPy's defaults (pgtol 1e-5, factr 1e7): rounds are cheap on the device, and the optimum then agrees
        # with a converged L-BFGS-B run to ~1e-7 instead of to SciPy's own stopping slack (~1e-4)
        self.lbfgs_options = dict(maxiter=1000, pgtol=1e-9, factr=10.0)
        self.lbfgs_options.update(lbfgs_options or {})
        self._device_gen = None
        self.last_stats = None

    def _device_lbfgs_target(self, duplicate_manager):
        """(engine, kind, par, lp_acquisition or None) when the local optimisation can run inside the library: L-BFGS on an
        all-continuous box without context variables or duplicate handling, on a fused acquisition with constant cost (or
        Local Penalization around one); None otherwise."""
        from .acquisitions import _FusedAcquisition, AcquisitionLP, _engine_of
        acq = self.acquisition
        if self.local_optimizer != 'device' or acq is None or self.optimizer_name != 'lbfgs' or duplicate_manager:
            return None
        if len(getattr(self.context_manager, 'context_index', [])) > 0:
            return None
        types = getattr(self.space, 'has_types', {})
        if not self.space.has_continuous() or any(v for t, v in types.items() if t != 'continuous') or self.space.has_constraints():
            return None
        if isinstance(acq, AcquisitionLP):
            if not acq._device_path_ok():
                return None
            return _engine_of(acq.model), acq.acq._kind, acq.acq._par(), acq
        if isinstance(acq, _FusedAcquisition) and _plain(acq):
            return _engine_of(acq.model), acq._kind, acq._par(), None
        return None

    def _optimize_on_device(self, target, anchor_points, f):
        eng, kind, par, lp = target
        bounds = np.asarray(self.context_manager.noncontext_bounds, dtype=np.float64)
        if lp is not None:
            eng.lp_set(lp.X_batch, lp.r_x0, lp.s_x0, lp.transform)
        try:
            X, fx, iters, nfev = eng.lbfgs(kind, par, np.atleast_2d(anchor_points), bounds, **self.lbfgs_options)
        finally:
            if lp is not None:
                eng.lp_set(None, None, None, None)
        # as apply_optimizer (optimizer.py:160-170): round to the design space, report f at the rounded point
        Xr = np.vstack([self.space.round_optimum(x[None, :]) for x in X])
        fr = np.asarray(f(Xr)).reshape(-1)
        best = int(np.argmin(fr))
        self.last_stats = {'device_calls': 1, 'rows': int(np.sum(nfev)), 'lbfgs_iterations': iters.tolist(), 'evaluations': nfev.tolist()}
        return np.atleast_2d(Xr[best]), np.atleast_2d(fr[best])

    def optimize(self, f=None, df=None, f_df=None, duplicate_manager=None):
        self.f, self.df, self.f_df = f, df, f_df
        self.optimizer = choose_optimizer(self.optimizer_name, self.context_manager.noncontext_bounds)
        if self.type_anchor_points_logic == max_objective_anchor_points_logic:
            if self.acquisition is not None and self.device_design_seed is not None:
                if self._device_gen is None or self._device_gen.acquisition is not self.acquisition:
                    self._device_gen = DeviceDesignAnchorPointsGenerator(self.space, self.acquisition, self.num_anchor_samples,
                                                                         self.device_design_seed)
                gen = self._device_gen
            elif self.acquisition is not None:
                gen = FusedAnchorPointsGenerator(self.space, random_design_type, self.acquisition, self.num_anchor_samples)
            else:
                gen = ObjectiveAnchorPointsGenerator(self.space, random_design_type, f)
        elif self.type_anchor_points_logic == thompson_sampling_anchor_points_logic:
            gen = FusedThompsonAnchorPointsGenerator(self.space, sobol_design_type, self.model)
        anchor_points = gen.get(duplicate_manager=duplicate_manager, context_manager=self.context_manager)

        target = self._device_lbfgs_target(duplicate_manager)
        if target is not None:
            return self._optimize_on_device(target, anchor_points, f)

        fns = {'f': f}
        if f_df is not None:
            fns['f_df'] = f_df
        rdv = Rendezvous(fns, len(anchor_points))
        bf = lambda x: rdv.call('f', x)
        bf_df = (lambda x: rdv.call('f_df', x)) if f_df is not None else None
        results = [None] * len(anchor_points)
        errors = []

        def work(i, a):
            try:
                results[i] = apply_optimizer(self.optimizer, a, f=bf, df=None, f_df=bf_df, duplicate_manager=duplicate_manager,
                                             context_manager=self.context_manager, space=self.space)
            except Exception as e:  # noqa: BLE001 - re-raised on the calling thread
                errors.append(e)
            finally:
                rdv.worker_done()

        threads = [threading.Thread(target=work, args=(i, a), daemon=True) for i, a in enumerate(anchor_points)]
        for t in threads:
            t.start()
        rdv.serve()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        self.last_stats = {'device_calls': rdv.calls, 'rows': rdv.rows}
        x_min, fx_min = min(results, key=lambda t: t[1])
        return x_min, fx_min
