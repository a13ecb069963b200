"""CPU oracle for the GP-posterior + acquisition hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``gpyopt_b200/`` may import this package; the
only legitimate importers are ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``, and there only as the checker
or as the CPU baseline being timed -- never as the thing shipped.

Contents
--------
``gpy_numpy``     NumPy/SciPy float64 restatement of the third-party package GPy's exact-GP
                  arithmetic (GPy>=1.8, un-pinned in /root/reference/setup.py:26; GPy is not
                  vendored in /root/reference and not installable here).  Restated from the
                  published GPy 1.9.x algorithm.
``gpyopt_numpy``  NumPy restatement of the GPyOpt-side formulas on the path
                  (GPModel.predict*, get_quantiles, EI/MPI/LCB(+_MCMC), LP, anchor top-k);
                  each function cites the reference file:line it follows.
``gpy_stub/GPy``  a stub ``GPy`` package (delegating to ``gpy_numpy``) so that the UNMODIFIED
                  reference ``/root/reference/GPyOpt`` imports in the build container; used by
                  ``tests/golden/make_golden.py`` to generate the committed golden vectors from
                  the reference's own acquisition / model-wrapper code.

Parity pin status: the GPyOpt-side formulas are pinned by the reference's own known-answer
tests (GPyOpt/testing/acquisitions_tests/*) and by golden vectors produced by running the
reference's code; the GPy-side arithmetic is pinned only by the reference's single real-GP
known answer (GPyOpt/testing/core_tests/test_cost.py:35-42, reproduced in
tests/test_oracle_pins.py).  For mean / variance / gradients at FIXED hyper-parameters the
reference holds no golden vector: at that boundary parity is "unpinned" beyond that test.
"""
