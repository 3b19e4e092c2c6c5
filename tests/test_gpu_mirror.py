"""The host-side mirror of the reference's call surface (decentralopf.jl_b200/admm.py) driven exactly like
/root/reference/src/opf_admm_decentral.jl:5-9, checked against the reference's committed traces, the thesis tables and
the oracle.  pytest -m gpu."""
import csv
import os

import numpy as np
import pytest

from tests.conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def traced(pkg):
    admm = pkg.ADMM(0.3, *pkg.cases.three_node())            # opf_admm_decentral.jl:5
    pkg.run(admm)                                            # :7
    return admm


def test_driver_script_surface(pkg, traced, oracle_mod):
    admm = traced
    import dopf_b200.admm as mod
    assert mod.admm is admm                                  # the helpers read the global `admm` like the reference
    assert admm.convergence.all and admm.iteration == 476    # stop rule of convergence.jl on the committed case
    # "at exit admm.iteration = k, there are k+1 dual entries and k results" 
    assert len(admm.lambdas) == len(admm.mues) == len(admm.rhos) == 477 and len(admm.results) == 476
    assert len(admm.convergence.lambda_res) == len(admm.convergence.mue_res) == len(admm.convergence.rho_res) == 475
    g = load_golden("TNS")
    gens, sto = admm.generators, admm.storages[0]
    worst = 0.0
    for k in range(476):
        r = admm.results[k]
        P = np.stack([r.unit_to_result[u].generation for u in gens])
        worst = max(worst, np.abs(P - g["P"][k]).max(), np.abs(r.unit_to_result[sto].discharge - g["D"][k][0]).max(),
                    np.abs(r.unit_to_result[sto].charge - g["C"][k][0]).max(),
                    np.abs(admm.lambdas[k] - g["lam"][k]).max(), np.abs(admm.mues[k] - g["mu"][k]).max(), np.abs(admm.rhos[k] - g["rho"][k]).max())
    assert worst < 2e-5
    # residual histories are what check_convergence! pushes (convergence.jl:5-12)
    np.testing.assert_array_equal(admm.convergence.lambda_res[-1], np.abs(admm.lambdas[-1] - admm.lambdas[-2]))
    assert max(admm.convergence.lambda_res[-1].max(), admm.convergence.mue_res[-1].max(), admm.convergence.rho_res[-1].max()) < 1e-3
    assert max(admm.convergence.lambda_res[-2].max(), admm.convergence.mue_res[-2].max(), admm.convergence.rho_res[-2].max()) >= 1e-3
    # np = get_nodal_price(admm.iteration)   (opf_admm_decentral.jl:9; Thesis Table 17, decentral column, all nodes;
    # signs flipped by construction, the central LP gives 36.6/82, 15.2/4, 30/30)
    np_ = pkg.get_nodal_price(admm.iteration)
    np.testing.assert_allclose(np_[0], [-36.5972, -81.9756], atol=2e-3)
    np.testing.assert_allclose(np_, -np.array([[36.6, 82.0], [15.2, 4.0], [30.0, 30.0]]), atol=0.05)
    ora = oracle_mod.OracleADMM(admm.problem, 0.3); ora.run(1000)
    np.testing.assert_allclose(np_, ora.nodal_price("prev"), rtol=0, atol=1e-6)
    # any iteration of the history, evaluated on the device from the host-kept duals
    k = 200
    expect = admm.lambdas[k - 1][None, :] + admm.ptdf.T @ (admm.mues[k - 1] + admm.rhos[k - 1])
    np.testing.assert_allclose(pkg.get_nodal_price(k), expect, atol=1e-10)
    with pytest.raises(IndexError):
        pkg.get_nodal_price(9999)
    # Result fields and the accessors of helpers/results.jl
    r = admm.results[-1]
    np.testing.assert_allclose(r.line_utilization, admm.ptdf @ r.injection, atol=1e-9)
    np.testing.assert_allclose(r.injection.sum(0), r.generation + r.discharge - r.charge - admm.total_demand, atol=1e-9)
    gen, dis, ch = pkg.get_results(476)
    assert gen is r.generation and dis is r.discharge and ch is r.charge
    assert pkg.get_unit_results(gens[2], 476) is r.unit_to_result[gens[2]].generation
    d_, c_ = pkg.get_unit_results(sto, 476)
    assert d_ is r.unit_to_result[sto].discharge and c_ is r.unit_to_result[sto].charge
    n1 = admm.nodes[0]
    np.testing.assert_allclose(pkg.get_node_results(476, n1)[0], r.unit_to_result[gens[0]].generation + r.unit_to_result[gens[3]].generation)
    aU, aK = pkg.get_average_slack_results(476)
    assert aU is r.avg_U and aK is r.avg_K
    assert abs(r.total_costs - 14035) / 14035 < 2e-4          # Thesis: central objective 14 035
    np.testing.assert_allclose(r.unit_to_result[sto].level, np.cumsum(r.unit_to_result[sto].charge - r.unit_to_result[sto].discharge), atol=1e-9)


def test_step_functions_publish_like_the_reference(pkg):
    admm = pkg.ADMM(0.3, *pkg.cases.three_node())
    assert admm.iteration == 1 and len(admm.results) == 0 and len(admm.lambdas) == 1
    z = pkg.get_unit_results(admm.generators[0], 0)
    assert z.shape == (2,) and not z.any()                       # helpers/results.jl:15-21: zeros before the first result
    pkg.optimize_all_subproblems(admm)
    assert len(admm.results) == 1 and len(admm.lambdas) == 1 and admm.iteration == 1
    with pytest.raises(RuntimeError):
        pkg.optimize_all_subproblems(admm)
    pkg.update_duals(admm)
    assert len(admm.lambdas) == 2 and admm.iteration == 1
    pkg.check_convergence(admm)
    assert admm.iteration == 2 and not admm.convergence.all and len(admm.convergence.lambda_res) == 0   # skipped at k = 1
    pkg.calculate_iteration(admm)
    assert admm.iteration == 3 and len(admm.convergence.lambda_res) == 1
    g = load_golden("TNS")
    np.testing.assert_allclose(admm.lambdas[2], g["lam"][2], atol=1e-6)


def test_unit_penalty_terms_and_private_slacks(pkg):
    """ResultGenerator/ResultStorage .penalty_term, .U, .K (subproblems.jl:89-102) and result.penalty_term
    (results.jl:73-76) against their definition (penalty_terms.jl:3-37) in numpy"""
    admm = pkg.ADMM(0.3, *pkg.cases.three_node())
    for _ in range(7):
        pkg.calculate_iteration(admm)
    prev, new = admm.results[-2], admm.results[-1]
    w, gamma = 10.0, 0.3
    kk = 2 * w + gamma
    tot = np.zeros((3, 2))
    units = admm.generators + admm.storages
    for u in units:
        ru, rp = new.unit_to_result[u], prev.unit_to_result[u]
        delta = (ru.generation - rp.generation) if hasattr(ru, "generation") else (ru.discharge - rp.discharge) - (ru.charge - rp.charge)
        p = admm.ptdf[:, admm.node_to_id[u.node] - 1][:, None]
        F, f = prev.line_utilization, admm.f_max[:, None]
        U = np.maximum(0, (2 * w * (f - F - p * delta) + gamma * prev.avg_U) / kk)
        K = np.maximum(0, (2 * w * (f + F + p * delta) + gamma * prev.avg_K) / kk)
        np.testing.assert_allclose(ru.U, U, atol=1e-9); np.testing.assert_allclose(ru.K, K, atol=1e-9)
        eb = (prev.injection.sum(0) + delta) ** 2
        up = ((F + p * delta + U - f) ** 2).sum(0); lo = ((K - F - p * delta - f) ** 2).sum(0)
        pt = ru.penalty_term
        np.testing.assert_allclose(pt.energy_balance, eb, rtol=1e-9, atol=1e-12); np.testing.assert_allclose(pt.upper_flow, up, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(pt.lower_flow, lo, rtol=1e-9, atol=1e-12)
        tot += np.stack([eb, up, lo])
    # avg_U is the mean of the private slacks (results.jl:83-84,110-112)
    np.testing.assert_allclose(new.avg_U, sum(new.unit_to_result[u].U for u in units) / len(units), atol=1e-9)
    t = new.penalty_term
    np.testing.assert_allclose(np.stack([t.energy_balance, t.upper_flow, t.lower_flow]), tot, rtol=1e-9, atol=1e-12)
    pkg.calculate_iteration(admm)
    with pytest.raises(RuntimeError):
        prev.unit_to_result[admm.generators[0]].U            # only the newest iterate can be evaluated on the device


def test_untraced_run_stays_on_device(pkg, traced):
    admm = pkg.ADMM(0.3, *pkg.cases.three_node(), trace=False)
    pkg.run(admm)
    assert admm.convergence.all and admm.iteration == 476 and len(admm.results) == 1 and len(admm.lambdas) == 2
    np.testing.assert_allclose(admm.lambdas[1], traced.lambdas[476], atol=1e-12)
    np.testing.assert_allclose(admm.lambdas[0], traced.lambdas[475], atol=1e-12)
    np.testing.assert_allclose(pkg.get_nodal_price(admm.iteration), pkg.get_nodal_price(476, traced), atol=1e-12)
    with pytest.raises(IndexError):
        pkg.get_nodal_price(100)
    with pytest.raises(RuntimeError):
        pkg.export_results(admm, "x")


def test_export_results_csv_format(pkg, traced, tmp_path):
    """helpers/output.jl:14-85: same files, columns, row order and values as the reference's results/TNS_*.csv
    (which run on to iteration 550; ours stops at the committed stop rule, iteration 476)"""
    pkg.export_results(traced, "TNS", parent_dir=str(tmp_path) + os.sep)
    g = load_golden("TNS")
    rows = list(csv.reader(open(tmp_path / "TNS_duals.csv")))
    assert rows[0] == ["iteration", "dual", "timestep", "line", "value"]
    assert rows[1] == ["1", "lambda", "1", "", "0.0"] and rows[2] == ["1", "lambda", "2", "", "0.0"]
    assert rows[3][:4] == ["2", "lambda", "1", ""] and abs(float(rows[3][4]) - (-24.53378055870387)) < 1e-6
    assert len(rows) == 1 + 476 * (2 + 6 + 6)
    assert rows[1 + 476 * 2][:4] == ["1", "rho", "1", "1"]                   # lambda block, then rho, then mue
    assert rows[1 + 476 * 8][:4] == ["1", "mue", "1", "1"]
    last_mue = rows[-1]
    assert last_mue[:4] == ["476", "mue", "2", "3"] and abs(float(last_mue[4]) - g["mu"][475][2, 1]) < 2e-5
    rows = list(csv.reader(open(tmp_path / "TNS_generators.csv")))
    assert rows[0] == ["iteration", "generator", "timestep", "generation"]
    assert rows[1][:3] == ["1", "pv", "1"] and abs(float(rows[1][3]) - 36.25469304166821) < 1e-6
    assert len(rows) == 1 + 4 * 476 * 2 and rows[1 + 476 * 2][:3] == ["1", "wind", "1"]
    rows = list(csv.reader(open(tmp_path / "TNS_storages.csv")))
    assert rows[0] == ["iteration", "storage", "timestep", "charge", "discharge"]
    assert rows[1][:3] == ["1", "battery", "1"] and abs(float(rows[1][3]) - 10.0) < 1e-6 and abs(float(rows[1][4])) < 1e-6
    assert len(rows) == 1 + 476 * 2
