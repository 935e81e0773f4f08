"""Pin the oracle against the reference's own golden values."""
import numpy as np

from oracle import solve as osolve
from tests import problems


def test_g1_readme_poisson_5x5():
    # tests/test_dicts_as_dofs_user_potential.py:62-63 (reference repo)
    p = problems.readme_poisson(5)
    prob = osolve.Problem(p["sets"], p["coords"], p["mask"], p["values"])
    assert int((~p["mask"]).sum()) == 20
    dofs, (steps, res, div) = osolve.damped_newton(prob, np.zeros(p["mask"].shape))
    assert steps == 1 and not div
    assert np.isclose(dofs.sum(), 1.9066412530282952, rtol=1e-12, atol=0)


def test_g2_cook_adaptive_load_stepping():
    # tests/test_user_elem_impl_diff_and_adaptive_load_step.py:128,153 (reference repo)
    p = problems.cook_g2()
    prob = osolve.Problem(p["sets"], p["coords"], p["mask"], p["values"])

    def mult(prob, m):
        prob.sets[1]["model"]["traction"] = np.array([0.0, m * p["q0"]])

    trace = []
    dofs, _ = osolve.adaptive_load_stepping(prob, np.zeros(p["mask"].shape), mult, newton_tol=1e-8, trace=trace)
    assert [t[1] for t in trace] == [6, 6, 5, 5, 4]
    assert np.allclose([t[0] for t in trace], [0.2, 0.414286, 0.643878, 0.906268, 1.0], atol=1e-6)
    assert np.isclose(dofs.ravel() @ dofs.ravel(), 19390.35027108, rtol=1e-10, atol=0)
