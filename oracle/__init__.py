"""CPU oracle for the Markov-GP inference hot path of AaltoML/BayesNewton.

TEST INFRASTRUCTURE ONLY.  Nothing under ``bayesnewton_b200/`` (the product)
may import this package: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only
as the checker / the timed CPU stand-in.

What it is: a NumPy (fp64, optionally ``np.longdouble``) restatement of the
reference's algorithm for the path SURVEY.md section 8 scopes --
``bayesnewton/ops.py:149-380`` (Kalman filter / RTS smoother, sequential and
associative-scan forms), the Matern state-space discretisation
(``kernels.py:123-382,1499-1616``), the cubature site updates
(``cubature.py:56-84,198-435``, ``likelihoods.py``, ``inference.py:21-428``)
and the energy terms (``basemodels.py:676-741``, ``utils.py:376-541``).

Pinning status.  The reference is pure Python on jax==0.4.14 / objax, neither
of which is installable in this image (no wheel, no network), so the reference
itself cannot be executed here and it ships no golden vectors.  The oracle is
therefore pinned by (i) the same cross-checks the reference's own test-suite
uses -- Markov (Kalman) path vs dense-GP path on the reference's parameter
grids (``tests/test_gp_vs_markovgp_{reg,class}.py``) and the closed-form
marginal likelihood (``tests/test_vs_exact_marg_lik.py``) -- restated with
fixed seeds in ``tests/test_oracle_*.py``, (ii) sequential-vs-scan agreement,
and (iii) a ``longdouble`` re-run.  The scan (``parallel=True``), EP/PL and
heteroscedastic paths have no test in the reference at all: for those rows the
status is "parity unpinned" (mathematical cross-checks only).
"""
