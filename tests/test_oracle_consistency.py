"""Internal consistency of the oracle: reduced form == literal formulation on random tiny grids,
and the generic QP solver against brute-force KKT checks (no GPU)."""
import numpy as np
import pytest


@pytest.mark.parametrize("seed,gamma,w", [(0, 0.1, 10.0), (1, 0.05, 10.0), (2, 0.02, 0.5), (3, 0.3 / 9, 1.0 / 9)])
def test_reduced_equals_literal(pkg, oracle_mod, seed, gamma, w):
    d = pkg.cases.synthetic_arrays(N=5, L=7, G=6, S=3, T=4, seed=seed, congest_frac=0.4)
    prob = pkg.Problem.from_arrays(d)
    o0 = oracle_mod.OracleADMM(prob, gamma, flow_weight=w)
    o1 = oracle_mod.OracleADMM(prob, gamma, flow_weight=w)
    for _ in range(30):
        o0.iterate(0); o1.iterate(1)
        for a, b in ((o0.P, o1.P), (o0.D, o1.D), (o0.C, o1.C), (o0.avgU, o1.avgU), (o0.avgK, o1.avgK), (o0.mu, o1.mu), (o0.rho, o1.rho), (o0.lam, o1.lam)):
            assert np.abs(a - b).max() < 1e-8 * max(1.0, np.abs(a).max())
    assert o0.qp_kkt_worst < 1e-6 and o1.qp_kkt_worst < 1e-6


def test_qp_solver_random(oracle_mod):
    rng = np.random.default_rng(0)
    for trial in range(40):
        n = int(rng.integers(2, 12)); m = int(rng.integers(1, 3 * n))
        A = rng.normal(size=(n, n)); G = A @ A.T + 0.1 * np.eye(n)
        g = rng.normal(size=n) * 3
        C = rng.normal(size=(m, n)); x0 = rng.normal(size=n)
        b = C @ x0 - rng.uniform(0, 1, size=m)          # x0 strictly feasible
        x, u, it, res = oracle_mod.qp_solve(G, g, C, b)
        assert it >= 0 and res < 1e-8
        # optimality by comparison with random feasible directions
        f = lambda z: 0.5 * z @ G @ z + g @ z
        for _ in range(20):
            z = x + 1e-3 * rng.normal(size=n)
            if np.all(C @ z >= b):
                assert f(z) >= f(x) - 1e-10


def test_qp_solver_detects_infeasible(oracle_mod):
    G = np.eye(2); g = np.zeros(2)
    C = np.array([[1.0, 0.0], [-1.0, 0.0]]); b = np.array([1.0, 0.0])   # x0 >= 1 and x0 <= 0
    x, u, it, res = oracle_mod.qp_solve(G, g, C, b)
    assert it == -1


def test_storage_level_edge_cases(pkg, oracle_mod):
    """empty horizon-like corner cases: T=1, a storage that cannot move (emax=0), zero demand."""
    d = pkg.cases.synthetic_arrays(N=4, L=5, G=3, S=2, T=1, seed=5)
    d["sto_emax"] = np.array([0.0, d["sto_emax"][1]])
    prob = pkg.Problem.from_arrays(d)
    o0 = oracle_mod.OracleADMM(prob, 0.05); o1 = oracle_mod.OracleADMM(prob, 0.05)
    for _ in range(10):
        o0.iterate(0); o1.iterate(1)
    assert np.abs(o0.E[0]).max() < 1e-9 and np.abs(o0.P - o1.P).max() < 1e-7
    # with emax = 0 the level constraint forces C_t = D_t; simultaneous charge/discharge is allowed
    assert np.abs(o0.C[0] - o0.D[0]).max() < 1e-9
