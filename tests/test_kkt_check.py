"""The KKT certificate (tests/kkt_check.py) itself, on the CPU: it must accept the oracle's iterates and reject perturbed ones."""
import numpy as np

from tests import kkt_check


def _snap(o):
    return dict(P=o.P.copy(), D=o.D.copy(), C=o.C.copy(), inj=o.inj.copy(), flow=o.flow.copy(), avgU=o.avgU.copy(), avgK=o.avgK.copy(),
                lam=o.lam.copy(), mu=o.mu.copy(), rho=o.rho.copy())


def test_certificate_accepts_oracle_and_rejects_perturbations(pkg, oracle_mod):
    for (dims, gs, ws, seed) in (((12, 18, 30, 8, 12), 0.3, 10.0, 3), ((20, 30, 40, 10, 24), 0.03, 1.0, 1)):
        N, L, G, S, T = dims
        prob = pkg.Problem.from_arrays(pkg.cases.synthetic_arrays(N=N, L=L, G=G, S=S, T=T, seed=seed, congest_frac=0.3))
        A = G + S
        gamma, w = gs / A, ws / A
        o = oracle_mod.OracleADMM(prob, gamma, flow_weight=w)
        for k in range(8):
            prev = _snap(o)
            o.iterate(0)
            scale = 1.0 + np.abs(prev["lam"]).max()
            assert kkt_check.generator_violation(prob, prev, o.P, gamma, w) < 1e-7 * scale
            assert kkt_check.storage_violation(prob, prev, o.D, o.C, gamma, w) < 1e-6 * scale
        # a feasible but suboptimal storage schedule (shift charge between two timesteps) must be rejected
        prev_ok = prev
        D2, C2 = o.D.copy(), o.C.copy()
        s = int(np.argmax((C2 > 1e-3).sum(1)))
        ts = np.where((C2[s] > 0.2) & (C2[s] < prob.sto_pmax[s] - 0.2))[0]
        if len(ts):
            C2[s, ts[0]] -= 0.1
            assert kkt_check.storage_violation(prob, prev_ok, D2, C2, gamma, w) > 1e-3
        P2 = o.P.copy()
        g = int(np.argmax(((P2 > 1) & (P2 < prob.gen_pmax[:, None] - 1)).sum(1)))
        tg = np.where((P2[g] > 1) & (P2[g] < prob.gen_pmax[g] - 1))[0]
        if len(tg):
            P2[g, tg[0]] += 0.5
            assert kkt_check.generator_violation(prob, prev_ok, P2, gamma, w) > 1e-2
