"""Parity of the CUDA path (through the C ABI, include/dopf.h) against the oracle, the reference's
golden traces and size-independent properties.  Needs a B200: pytest -m gpu.

Tolerances (north_star): iterates after a fixed number of iterations within 1e-6 relative of the
reference path; the golden traces themselves carry Gurobi's tolerance (<= 5.7e-6 abs).
"""
import numpy as np
import pytest

from tests.conftest import GOLDEN, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def DeviceADMM(pkg):
    from dopf_b200.device import DeviceADMM as D
    return D


def _rel(a, b):
    return np.abs(a - b).max() / max(1.0, np.abs(b).max()) if a.size else 0.0


def _compare(dev, ora, tol):
    it = dev.get_iterate(); lam, mu, rho = dev.get_duals(0)
    pairs = dict(P=(it["P"], ora.P), D=(it["D"], ora.D), C=(it["C"], ora.C), E=(it["E"], ora.E), inj=(it["injection"], ora.inj),
                 flow=(it["flow"], ora.flow), avgU=(it["avgU"], ora.avgU), avgK=(it["avgK"], ora.avgK), lam=(lam, ora.lam), mu=(mu, ora.mu), rho=(rho, ora.rho))
    flips = ((mu == 0) != (ora.mu == 0)).sum() + ((rho == 0) != (ora.rho == 0)).sum()
    assert flips == 0, "slack-mask flip (discontinuous dual mask) - reported separately from numeric drift"
    for k, (a, b) in pairs.items():
        assert _rel(a, b) < tol, (k, _rel(a, b))


@pytest.mark.parametrize("name", GOLDEN)
def test_golden_traces_through_c_abi(pkg, DeviceADMM, name):
    """reference order of agents (pv, wind, coal, gas): exercises the library's node sort/permutation"""
    prob = pkg.Problem.from_structs(*pkg.cases.three_node())
    g = load_golden(name)
    dev = DeviceADMM(prob, gamma=float(g["gamma"]), flow_weight=float(g["flow_weight"]), device=0)
    worst, stop = 0.0, None
    for k in range(g["P"].shape[0]):
        lam, mu, rho = dev.get_duals(0)
        worst = max(worst, np.abs(lam - g["lam"][k]).max(), np.abs(mu - g["mu"][k]).max(), np.abs(rho - g["rho"][k]).max())
        st = dev.step(1)
        if st.converged:
            stop = k + 1
            break
        it = dev.get_iterate(("P", "D", "C"))
        worst = max(worst, np.abs(it["P"] - g["P"][k]).max(), np.abs(it["D"] - g["D"][k]).max(), np.abs(it["C"] - g["C"][k]).max())
    assert worst < 2e-5
    assert stop == (476 if name == "TNS" else None)
    if name == "TNS":
        assert dev.iteration == 476 and dev.step(5).iterations_done == 476     # run! is a no-op once converged
        np.testing.assert_allclose(dev.nodal_price(1)[0], [-36.5972, -81.9756], atol=2e-3)   # Thesis Table 17
        assert abs(dev.total_costs() - 14035) / 14035 < 2e-4                                 # Thesis: central objective 14 035


def test_run_to_convergence_in_one_call(pkg, DeviceADMM, oracle_mod):
    prob = pkg.Problem.from_structs(*pkg.cases.three_node())
    dev = DeviceADMM(prob, gamma=0.3, device=0)
    st = dev.step(100000)                       # run!(admm): stops on the device-side convergence flag
    assert st.converged and st.iteration == 476 and st.iterations_done == 476
    ora = oracle_mod.OracleADMM(prob, 0.3); ora.run(1000)
    _compare(dev, ora, 1e-6)


@pytest.mark.parametrize("dims,gamma,w,iters,seed", [
    ((12, 18, 30, 8, 6), None, None, 40, 3), ((12, 18, 30, 8, 6), 0.02, 10.0, 25, 3), ((5, 7, 6, 3, 4), 0.1, 10.0, 30, 1),
    ((40, 60, 200, 40, 24), None, None, 20, 1), ((118, 186, 1000, 200, 24), None, None, 25, 3), ((33, 50, 64, 16, 7), None, None, 20, 9)])
def test_iterates_match_oracle(pkg, DeviceADMM, oracle_mod, dims, gamma, w, iters, seed):
    N, L, G, S, T = dims
    d = pkg.cases.synthetic_arrays(N=N, L=L, G=G, S=S, T=T, seed=seed)
    prob = pkg.Problem.from_arrays(d)
    A = G + S
    gamma = gamma or 0.3 / A; w = w or 1.0 / A
    dev = DeviceADMM(prob, gamma=gamma, flow_weight=w, device=0, hinge_capacity=64)
    ora = oracle_mod.OracleADMM(prob, gamma, flow_weight=w)
    for k in range(iters):
        dev.step(1); ora.iterate(0)
        _compare(dev, ora, 1e-6)
    assert dev.status.gen_corrected > 0         # the exact-correction pass ran


def test_batched_steps_equal_single_steps_and_graph_equals_direct(pkg, DeviceADMM):
    d = pkg.cases.synthetic_arrays(N=40, L=60, G=200, S=40, T=24, seed=2)
    prob = pkg.Problem.from_arrays(d); A = 240
    a = DeviceADMM(prob, gamma=0.3 / A, flow_weight=1.0 / A, device=0)
    b = DeviceADMM(prob, gamma=0.3 / A, flow_weight=1.0 / A, device=0, use_graph=False)
    a.step(17)
    for _ in range(17):
        b.step(1)
    ia, ib = a.get_iterate(), b.get_iterate()
    for k in ia:
        assert _rel(ia[k], ib[k]) < 1e-9, k


def test_unsorted_agents_and_edge_cases(pkg, DeviceADMM, oracle_mod):
    """agents given in random node order; odd T (scalar generator kernel); node without agents;
    a storage with zero capacity; a generator with zero capacity."""
    d = pkg.cases.synthetic_arrays(N=9, L=12, G=14, S=5, T=5, seed=4)
    rng = np.random.default_rng(0)
    pg, ps = rng.permutation(14), rng.permutation(5)
    for k in ("gen_mc", "gen_pmax", "gen_node"):
        d[k] = d[k][pg]
    for k in ("sto_mc", "sto_pmax", "sto_emax", "sto_node"):
        d[k] = d[k][ps]
    d["gen_node"][d["gen_node"] == 3] = 4          # node 3 hosts no generator
    d["sto_node"][d["sto_node"] == 3] = 5
    d["sto_emax"][0] = 0.0; d["gen_pmax"][1] = 0.0
    prob = pkg.Problem.from_arrays(d)
    dev = DeviceADMM(prob, gamma=0.02, flow_weight=0.5, device=0); ora = oracle_mod.OracleADMM(prob, 0.02, flow_weight=0.5)
    for _ in range(30):
        dev.step(1); ora.iterate(0)
        _compare(dev, ora, 1e-6)


def test_agents_on_a_node_subrange(pkg, DeviceADMM, oracle_mod):
    """all agents sit on the upper nodes: the PTDF^T product is then restricted to the row tiles of that node range
    (what every rank of the agent-partitioned mode does); many storages per node share their hinge lists"""
    d = pkg.cases.synthetic_arrays(N=150, L=220, G=300, S=90, T=12, seed=6)
    d["gen_node"] = (70 + d["gen_node"] % 80).astype(d["gen_node"].dtype)
    d["sto_node"] = (128 + d["sto_node"] % 20).astype(d["sto_node"].dtype)
    prob = pkg.Problem.from_arrays(d); A = 390
    dev = DeviceADMM(prob, gamma=0.3 / A, flow_weight=1.0 / A, device=0, hinge_capacity=64)
    ora = oracle_mod.OracleADMM(prob, 0.3 / A, flow_weight=1.0 / A)
    for _ in range(20):
        dev.step(1); ora.iterate(0)
        _compare(dev, ora, 1e-6)
    assert dev.status.gen_corrected > 0 and dev.status.sto_corrected > 0


def test_generators_only_and_storages_only(pkg, DeviceADMM, oracle_mod):
    for G, S in ((8, 0), (0, 6)):
        d = pkg.cases.synthetic_arrays(N=6, L=8, G=max(G, 1), S=max(S, 1), T=6, seed=G + S)
        if G == 0:
            for k in ("gen_mc", "gen_pmax", "gen_node"):
                d[k] = d[k][:0]
            d["G"] = 0
        if S == 0:
            for k in ("sto_mc", "sto_pmax", "sto_emax", "sto_node"):
                d[k] = d[k][:0]
            d["S"] = 0
        prob = pkg.Problem.from_arrays(d)
        dev = DeviceADMM(prob, gamma=0.03, flow_weight=1.0, device=0); ora = oracle_mod.OracleADMM(prob, 0.03, flow_weight=1.0)
        for _ in range(15):
            dev.step(1); ora.iterate(0)
        _compare(dev, ora, 1e-6)


def test_set_state_resumes_a_trajectory(pkg, DeviceADMM):
    d = pkg.cases.synthetic_arrays(N=40, L=60, G=200, S=40, T=24, seed=5)
    prob = pkg.Problem.from_arrays(d); A = 240
    a = DeviceADMM(prob, gamma=0.3 / A, flow_weight=1.0 / A, device=0)
    a.step(12)
    it = a.get_iterate(); lam, mu, rho = a.get_duals(0)
    b = DeviceADMM(prob, gamma=0.3 / A, flow_weight=1.0 / A, device=0)
    b.set_state(a.iteration, P=it["P"], D=it["D"], C_=it["C"], avgU=it["avgU"], avgK=it["avgK"], lam=lam, mu=mu, rho=rho)
    jb = b.get_iterate()
    assert _rel(jb["injection"], it["injection"]) < 1e-12 and _rel(jb["flow"], it["flow"]) < 1e-10 and _rel(jb["E"], it["E"]) < 1e-12
    a.step(8); b.step(8)
    ia, ib = a.get_iterate(), b.get_iterate()
    for k in ia:
        assert _rel(ia[k], ib[k]) < 1e-8, k
    assert a.iteration == b.iteration


def test_hinge_capacity_overflow_is_loud(pkg, DeviceADMM):
    """a too small hinge list capacity must stop the run with an error, never silently truncate"""
    from dopf_b200.device import DopfError
    raised = 0
    for seed in range(8):
        d = pkg.cases.synthetic_arrays(N=12, L=18, G=30, S=8, T=6, seed=seed, congest_frac=0.5)
        prob = pkg.Problem.from_arrays(d)
        dev = DeviceADMM(prob, gamma=0.02, flow_weight=10.0, device=0, hinge_capacity=1)
        try:
            dev.step(40)
        except DopfError as e:
            assert "hinge list capacity" in str(e)
            raised += 1
    assert raised > 0


@pytest.mark.parametrize("dims", [(2000, 3000, 20000, 5000, 96)])
def test_full_size_properties(pkg, DeviceADMM, dims):
    """BASELINE configs[2] shape: properties that need no oracle."""
    N, L, G, S, T = dims
    d = pkg.cases.synthetic_arrays(N=N, L=L, G=G, S=S, T=T, seed=0)
    prob = pkg.Problem.from_arrays(d); A = G + S
    gamma, w = 0.3 / A, 1.0 / A
    dev = DeviceADMM(prob, gamma=gamma, flow_weight=w, device=0, hinge_capacity=64)
    dev.step(6)
    lam0, mu0, rho0 = dev.get_duals(0)
    dev.step(1)
    it = dev.get_iterate(); lam1, mu1, rho1 = dev.get_duals(0); lamp, mup, rhop = dev.get_duals(1)
    assert np.array_equal(lamp, lam0) and np.array_equal(mup, mu0)
    # boxes and levels (subproblems.jl:26,114-116,150-156)
    assert it["P"].min() >= 0 and (it["P"] <= prob.gen_pmax[:, None] + 1e-12).all()
    assert it["D"].min() >= 0 and it["C"].min() >= 0 and (it["D"] <= prob.sto_pmax[:, None] + 1e-12).all() and (it["C"] <= prob.sto_pmax[:, None] + 1e-12).all()
    np.testing.assert_allclose(it["E"], np.cumsum(it["C"] - it["D"], axis=1), atol=1e-9)
    assert it["E"].min() > -1e-7 and (it["E"] <= prob.sto_emax[:, None] + 1e-7).all()
    # aggregation (results.jl:64,88-106,114)
    inj = -prob.demand.copy()
    np.add.at(inj, prob.gen_node, it["P"]); np.add.at(inj, prob.sto_node, it["D"] - it["C"])
    np.testing.assert_allclose(it["injection"], inj, atol=1e-8)
    np.testing.assert_allclose(it["flow"], prob.ptdf @ it["injection"], rtol=0, atol=1e-7 * np.abs(it["flow"]).max())
    # dual updates (update_duals.jl:7-39)
    np.testing.assert_allclose(lam1, lam0 + gamma * it["injection"].sum(0), atol=1e-10)
    mu_ref = (mu0 + gamma * (it["flow"] + it["avgU"] - prob.fmax[:, None])) * (it["avgU"] <= 1e-2)
    rho_ref = (rho0 + gamma * (it["avgK"] - it["flow"] - prob.fmax[:, None])) * (it["avgK"] <= 1e-2)
    np.testing.assert_allclose(mu1, mu_ref, atol=1e-10); np.testing.assert_allclose(rho1, rho_ref, atol=1e-10)
    assert it["avgU"].min() >= 0 and it["avgK"].min() >= 0
    st = dev.status
    assert abs(st.res_lambda - np.abs(lam1 - lam0).max()) < 1e-12 and abs(st.res_mue - np.abs(mu1 - mu0).max()) < 1e-12


def test_average_slacks_match_per_agent_definition_at_scale(pkg, DeviceADMM):
    """avg_U = mean over agents of U*(delta_i) (results.jl:83-84,110-112) recomputed in numpy from the
    previous and the new iterate on the 118-node case."""
    N, L, G, S, T = 118, 186, 1000, 200, 24
    d = pkg.cases.synthetic_arrays(N=N, L=L, G=G, S=S, T=T, seed=2)
    prob = pkg.Problem.from_arrays(d); A = G + S
    gamma, w = 0.3 / A, 1.0 / A
    dev = DeviceADMM(prob, gamma=gamma, flow_weight=w, device=0, hinge_capacity=64)
    dev.step(9)
    a = dev.get_iterate()
    dev.step(1)
    b = dev.get_iterate()
    kk = 2 * w + gamma
    nodes = np.concatenate([prob.gen_node, prob.sto_node])
    delta = np.concatenate([b["P"] - a["P"], (b["D"] - a["D"]) - (b["C"] - a["C"])])          # [A,T]
    ap = prob.fmax[:, None] - a["flow"]; am = prob.fmax[:, None] + a["flow"]
    U = np.zeros((L, T)); K = np.zeros((L, T))
    for i in range(A):
        pd = prob.ptdf[:, nodes[i]][:, None] * delta[i][None, :]
        U += np.maximum(0, (2 * w * (ap - pd) + gamma * a["avgU"]) / kk)
        K += np.maximum(0, (2 * w * (am + pd) + gamma * a["avgK"]) / kk)
    np.testing.assert_allclose(b["avgU"], U / A, rtol=0, atol=1e-9 * max(1, (U / A).max()))
    np.testing.assert_allclose(b["avgK"], K / A, rtol=0, atol=1e-9 * max(1, (K / A).max()))
