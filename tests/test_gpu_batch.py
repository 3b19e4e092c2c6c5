"""Scenario batches (BASELINE configs[3], SURVEY.md 8(e) "scenario"): C independent problems on one grid run as one
device problem.  The batched run must equal the loop over single-scenario handles - bit for bit when both use the same
split-K (summation order) of the PTDF products - including every scenario's own stop iteration.  pytest -m gpu."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run_batch_and_singles(pkg, dims, C, iters, cfg, seed=0, eps=1e-3):
    from dopf_b200.device import DeviceADMM
    N, L, G, S, T = dims
    d = pkg.cases.synthetic_scenarios(N=N, L=L, G=G, S=S, T=T, n_scen=C, seed=seed)
    prob = pkg.Problem.from_arrays(d)
    bat = DeviceADMM(prob, device=0, eps=eps, **cfg)
    bat.step(iters)
    singles = []
    for c in range(C):
        one = DeviceADMM(prob.scenario(c), device=0, eps=eps, **cfg)
        one.step(iters)
        singles.append(one)
    return prob, bat, singles


def test_batched_equals_loop_over_single_scenarios_bitwise(pkg):
    dims, C = (40, 60, 200, 40, 24), 5
    A = dims[2] + dims[3]
    cfg = dict(gamma=0.3 / A, flow_weight=10.0 / A, hinge_capacity=64, gemm_ksplit=1)
    prob, bat, singles = _run_batch_and_singles(pkg, dims, C, 25, cfg)
    it = bat.get_iterate(); lam, mu, rho = bat.get_duals(0); npz = bat.nodal_price(0); tc = bat.total_costs()
    assert it["P"].shape == (C, dims[2], dims[4]) and it["flow"].shape == (C, dims[1], dims[4]) and lam.shape == (C, dims[4])
    fixes = 0
    for c, one in enumerate(singles):
        oi = one.get_iterate(); ol, om, orr = one.get_duals(0)
        for k in it:
            assert np.array_equal(it[k][c], oi[k]), (c, k, np.abs(it[k][c] - oi[k]).max())
        assert np.array_equal(lam[c], ol) and np.array_equal(mu[c], om) and np.array_equal(rho[c], orr)
        assert np.array_equal(npz[c], one.nodal_price(0))
        assert abs(tc[c] - one.total_costs()) <= 1e-9 * abs(tc[c])
        fixes += one.status.gen_corrected
    assert bat.status.gen_corrected == fixes > 0          # the hinge correction ran, for exactly the same agents
    # scenarios differ (own demand / cost draws)
    assert not np.array_equal(it["P"][0], it["P"][1])


def test_every_scenario_stops_on_its_own(pkg):
    """scenarios of the reference's three-node case (literal gamma 0.3, weight 10, eps 1e-3) with scaled demand: scenario 0 is
    the committed case and stops at iteration 476 like the reference's trace; the others stop at their own iterations.
    Each must freeze at exactly the iterate and iteration count its single run stops at, while the others continue."""
    from dopf_b200.device import DeviceADMM
    base = pkg.Problem.from_structs(*pkg.cases.three_node())
    scale = [1.0, 0.95, 1.04, 0.9, 1.0]
    C = len(scale)
    tile = lambda a: np.tile(a, (C, 1))
    prob = pkg.Problem.from_arrays(dict(N=3, L=3, T=2, G=4, S=1, n_scen=C, ptdf=base.ptdf, fmax=base.fmax,
                                        demand=np.stack([base.demand * f for f in scale]), gen_mc=tile(base.gen_mc), gen_pmax=tile(base.gen_pmax),
                                        gen_node=base.gen_node, sto_mc=tile(base.sto_mc), sto_pmax=tile(base.sto_pmax), sto_emax=tile(base.sto_emax),
                                        sto_node=base.sto_node))
    bat = DeviceADMM(prob, device=0, gamma=0.3, gemm_ksplit=1)
    st = bat.step(5000)                                  # run!: returns when ALL scenarios have converged
    singles = []
    for c in range(C):
        one = DeviceADMM(prob.scenario(c), device=0, gamma=0.3, gemm_ksplit=1)
        one.step(5000)
        singles.append(one)
    its, conv, res = bat.scenario_status()
    one_its = np.array([o.status.iteration for o in singles]); one_conv = np.array([o.status.converged for o in singles])
    assert conv.all() and one_conv.all() and st.converged
    assert np.array_equal(its, one_its), (its, one_its)
    assert its[0] == 476 and its[4] == 476 and len(set(its.tolist())) >= 3, its      # they really stop at different iterations
    assert st.iterations_done == its.max()
    it = bat.get_iterate(); lam, mu, rho = bat.get_duals(0); lamp, mup, rhop = bat.get_duals(1)
    for c, one in enumerate(singles):
        oi = one.get_iterate(); ol, om, orr = one.get_duals(0)
        for k in it:
            assert np.array_equal(it[k][c], oi[k]), (c, k)
        assert np.array_equal(lam[c], ol) and np.array_equal(mu[c], om) and np.array_equal(rho[c], orr)
        np.testing.assert_array_equal(res[c], [one.status.res_lambda, one.status.res_mue, one.status.res_rho])
    np.testing.assert_allclose(bat.nodal_price(0)[0][0], [-36.5972, -81.9756], atol=2e-3)      # Thesis Table 17 for the committed case


def test_batch_against_oracle(pkg, oracle_mod):
    dims, C = (30, 45, 60, 12, 12), 3
    from dopf_b200.device import DeviceADMM
    N, L, G, S, T = dims
    A = G + S
    d = pkg.cases.synthetic_scenarios(N=N, L=L, G=G, S=S, T=T, n_scen=C, seed=2)
    prob = pkg.Problem.from_arrays(d)
    bat = DeviceADMM(prob, device=0, gamma=0.3 / A, flow_weight=10.0 / A, hinge_capacity=64)
    oras = [oracle_mod.OracleADMM(prob.scenario(c), 0.3 / A, flow_weight=10.0 / A) for c in range(C)]
    for _ in range(12):
        bat.step(1)
        for o in oras:
            o.iterate(0)
    it = bat.get_iterate(); lam, mu, rho = bat.get_duals(0)
    for c, o in enumerate(oras):
        for a, b in ((it["P"][c], o.P), (it["D"][c], o.D), (it["C"][c], o.C), (it["E"][c], o.E), (it["injection"][c], o.inj), (it["flow"][c], o.flow),
                     (it["avgU"][c], o.avgU), (it["avgK"][c], o.avgK), (lam[c], o.lam), (mu[c], o.mu), (rho[c], o.rho)):
            assert np.abs(a - b).max() <= 1e-6 * max(1.0, np.abs(b).max())


def test_large_batch_properties(pkg):
    """128 scenarios of the 118-node / 24-period case (one GPU's share of BASELINE configs[3]): aggregation and dual
    update identities per scenario"""
    from dopf_b200.device import DeviceADMM
    N, L, G, S, T, C = 118, 186, 1000, 200, 24, 128
    A = G + S
    d = pkg.cases.synthetic_scenarios(N=N, L=L, G=G, S=S, T=T, n_scen=C, seed=0)
    prob = pkg.Problem.from_arrays(d)
    gamma, w = 0.03 / A, 1.0 / A
    bat = DeviceADMM(prob, device=0, gamma=gamma, flow_weight=w, hinge_capacity=64)
    bat.step(8)
    lam0, mu0, rho0 = bat.get_duals(0)
    bat.step(1)
    it = bat.get_iterate(); lam1, mu1, rho1 = bat.get_duals(0)
    for c in (0, 17, 127):
        inj = -prob.demand[c].copy()
        np.add.at(inj, prob.gen_node, it["P"][c]); np.add.at(inj, prob.sto_node, it["D"][c] - it["C"][c])
        np.testing.assert_allclose(it["injection"][c], inj, atol=1e-8)
        np.testing.assert_allclose(it["flow"][c], prob.ptdf @ it["injection"][c], rtol=0, atol=1e-7 * np.abs(it["flow"][c]).max())
        np.testing.assert_allclose(lam1[c], lam0[c] + gamma * it["injection"][c].sum(0), atol=1e-10)
        mu_ref = (mu0[c] + gamma * (it["flow"][c] + it["avgU"][c] - prob.fmax[:, None])) * (it["avgU"][c] <= 1e-2)
        np.testing.assert_allclose(mu1[c], mu_ref, atol=1e-10)
        np.testing.assert_allclose(it["E"][c], np.cumsum(it["C"][c] - it["D"][c], axis=1), atol=1e-9)
        assert it["P"][c].min() >= 0 and (it["P"][c] <= prob.gen_pmax[c][:, None] + 1e-12).all()
    its, conv, _ = bat.scenario_status()
    assert (its == 10).all() and not conv.any()
