"""Long horizons (BASELINE configs[4] groundwork: T up to 8760) and the benchmark size, verified with the KKT certificate
of tests/kkt_check.py - the dense-QP oracle needs minutes per storage beyond T ~ 500 (O(T^3)), the certificate is exact at
any size (every agent problem is strictly convex: KKT <=> the reference's unique solution).  pytest -m gpu."""
import numpy as np
import pytest

from tests import kkt_check

pytestmark = pytest.mark.gpu


def _certify(pkg, dims, gs, ws, iters, seed=0, sample=None, hcap=64, **kw):
    from dopf_b200.device import DeviceADMM
    N, L, G, S, T = dims
    prob = pkg.Problem.from_arrays(pkg.cases.synthetic_arrays(N=N, L=L, G=G, S=S, T=T, seed=seed, **kw))
    A = G + S
    gamma, w = gs / A, ws / A
    dev = DeviceADMM(prob, gamma=gamma, flow_weight=w, device=0, hinge_capacity=hcap)
    rng = np.random.default_rng(0)
    for k in range(iters):
        prev = kkt_check.snapshot(dev)
        dev.step(1)
        new = dev.get_iterate()
        sub, p2, n2 = prob, prev, new
        if sample:                                  # certificate on a random sample of the agents (the check is O(L*T) per agent in numpy)
            gi = rng.choice(G, min(sample[0], G), replace=False); si = rng.choice(S, min(sample[1], S), replace=False)
            sub = pkg.Problem(N, L, T, len(gi), len(si), prob.ptdf, prob.fmax, prob.demand, prob.gen_mc[gi], prob.gen_pmax[gi], prob.gen_node[gi],
                              prob.sto_mc[si], prob.sto_pmax[si], prob.sto_emax[si], prob.sto_node[si])
            p2 = dict(prev, P=prev["P"][gi], D=prev["D"][si], C=prev["C"][si])
            n2 = dict(P=new["P"][gi], D=new["D"][si], C=new["C"][si])
        scale = 1.0 + np.abs(prev["lam"]).max() + np.abs(prev["mu"]).max()
        vg = kkt_check.generator_violation(sub, p2, n2["P"], gamma, w)
        vs = kkt_check.storage_violation(sub, p2, n2["D"], n2["C"], gamma, w)
        assert vg < 1e-7 * scale, ("generator KKT", k, vg)
        assert vs < 1e-6 * scale, ("storage KKT", k, vs)
        np.testing.assert_allclose(new["E"], np.cumsum(new["C"] - new["D"], axis=1), atol=1e-8)
    return dev


@pytest.mark.parametrize("gs,ws", [(0.3, 10.0)])
def test_T1024_small_grid(pkg, gs, ws):
    dev = _certify(pkg, (20, 30, 40, 4, 1024), gs, ws, iters=4, seed=1, congest_frac=0.3)
    assert dev.status.gen_corrected > 0


def test_T8760_hourly_year(pkg):
    """one year of hourly periods (the horizon of BASELINE configs[4]) on a small grid"""
    dev = _certify(pkg, (10, 14, 12, 2, 8760), 0.3, 10.0, iters=2, seed=2, congest_frac=0.3)
    assert dev.status.iterations_done == 2


def test_benchmark_size_sample_certificate(pkg):
    """the benchmarked case itself (2000 nodes / 3000 lines / 80k generators + 20k storages / 96 periods, bench.py's default
    parameters): KKT certificate on a random sample of agents in the cold-start transient the driver times"""
    _certify(pkg, (2000, 3000, 80000, 20000, 96), 0.03, 1.0, iters=3, sample=(100, 40))
