"""CPU oracle for the AutoPDEx hot path (sparse assembly -> Newton linear solve).

TEST INFRASTRUCTURE ONLY.  Nothing in ``autopdex_b200`` may import this package;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use it, and there only as the checker / the timed CPU
baseline, never as a compute path of the product.

This is a NumPy/SciPy *restatement* of the reference algorithm (JAX is not
installable in this image, so the reference itself cannot be executed here; see
SURVEY.md section 8c).  Every function cites the reference file:line it follows
(paths relative to the AutoPDEx v1.1.4 tree).

Parity pins (tests/test_oracle_golden.py):
  * G1  tests/test_dicts_as_dofs_user_potential.py:62-63   sum(phi) = 1.9066412530282952
  * G2  tests/test_user_elem_impl_diff_and_adaptive_load_step.py:153   u.u = 19390.35027108
  * outputs of the UNMODIFIED reference modules executed on a NumPy stand-in for JAX
    (tests/golden/fakejax.py, generator tests/golden/make_reference_fixtures.py), committed as
    tests/golden/reference_fixtures.npz and checked in tests/test_reference_fixtures.py
The SciPy half of the reference path (coo->csr duplicate summing, row/column
deletion, spsolve) is executed as-is with the same SciPy calls.
"""
