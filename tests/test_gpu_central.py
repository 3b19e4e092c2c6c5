"""The reference's own acceptance criterion beyond three_node: converged decentral ADMM vs the central LP
(/root/reference/src/opf_central_reference.jl:16-57, restated with scipy/HiGHS in tests/central_lp.py), Thesis section 4.3:
differences "in the per mille range" (max rel. 1.84e-4 on the three-node system).

Only small synthetic systems reach the reference's stop rule: from a few dozen agents on, the Jacobi update with linear
generator costs settles into a persistent oscillation of the duals (residuals ~3e-2 after 400 000 iterations on a 30-node /
72-agent case, scripts/converge_probe.py), so the criterion cannot be evaluated at the benchmark
sizes - a property of the reference algorithm, identical in the oracle.  pytest -m gpu."""
import numpy as np
import pytest

from tests import central_lp

pytestmark = pytest.mark.gpu


def _check(pkg, oracle_mod, prob, gamma, expect_iteration=None):
    from dopf_b200.device import DeviceADMM
    lp = central_lp.solve(prob)
    assert lp["status"] == 0
    dev = DeviceADMM(prob, gamma=gamma, device=0)            # literal flow weight 10, eps 1e-3
    st = dev.step(20000)
    assert st.converged
    ora = oracle_mod.OracleADMM(prob, gamma); ora.run(20000)
    assert ora.converged and ora.iteration == st.iteration   # time to tolerance: the same stop iteration as the CPU restatement
    if expect_iteration:
        assert st.iteration == expect_iteration
    it = dev.get_iterate()
    assert np.abs(it["P"] - lp["P"]).max() / np.abs(lp["P"]).max() < 5e-4
    assert abs(dev.total_costs() - lp["objective"]) / lp["objective"] < 5e-4
    assert np.abs(it["flow"] - lp["flow"]).max() < 5e-3 * max(1.0, np.abs(lp["flow"]).max())
    assert np.abs(it["injection"].sum(0)).max() < 0.05       # energy balance of the converged point
    # system price: lambda converges to minus the LP's energy-balance dual (sign flipped by construction, Thesis Table 16)
    lam = dev.get_duals(0)[0]
    return st, lam, lp


def test_three_node_vs_central_lp(pkg, oracle_mod):
    prob = pkg.Problem.from_structs(*pkg.cases.three_node())
    st, lam, lp = _check(pkg, oracle_mod, prob, 0.3, expect_iteration=476)
    assert abs(lp["objective"] - 14035) < 1e-6               # Thesis: central objective
    np.testing.assert_allclose(-lam, [30.0, 30.0], atol=2e-3)  # Thesis Table 16/17: system price 30 / 30


@pytest.mark.parametrize("seed,dims,gamma", [(17, (6, 8, 5, 1, 3), 0.15), (20, (7, 7, 6, 1, 4), 0.05)])
def test_small_synthetic_systems_vs_central_lp(pkg, oracle_mod, seed, dims, gamma):
    N, L, G, S, T = dims
    prob = pkg.Problem.from_arrays(pkg.cases.synthetic_arrays(N=N, L=L, G=G, S=S, T=T, seed=seed, congest_frac=0.0))
    _check(pkg, oracle_mod, prob, gamma)
