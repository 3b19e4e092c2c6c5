"""Size-independent optimality certificate of one ADMM iteration (test infrastructure).

Every agent problem of the reference (subproblems.jl:19-207 + penalty_terms.jl) is strictly convex, so an iterate is THE
solution iff it satisfies the KKT conditions.  This module checks them in numpy for a new iterate given the state the
iteration started from - independently of the CUDA solvers and of the oracle's QP solver - so horizons far beyond what
the dense-QP oracle can afford (T = 1024 ... 8760) are still verified exactly.

generator (1-D): g'(delta) = mc + pi + gamma(Sbar + delta) + prox*delta + h(delta) is 0 in the interior, >= 0 at P = 0, <= 0 at P = pmax.
storage: there must be a level-multiplier path eta_t with eta_{T+1} = 0, constant while 0 < E_t < emax, not rising across
a full level (E_t = emax), not falling across an empty one (E_t = 0), and stationarity of D_t, C_t with their box multipliers.
h is the slack-eliminated flow penalty derivative of SURVEY.md A.2, evaluated with all 2L hinges.
"""
import numpy as np


def _hinge_force(ptdf_col, delta, prev, fmax, gamma, w):
    """h_nt(delta) for one node: [T] (delta [T]); prev: flow, avgU, avgK [L,T]"""
    p = ptdf_col[:, None]
    kk = 2 * w + gamma
    ap = fmax[:, None] - prev["flow"]; am = fmax[:, None] + prev["flow"]
    pd = p * delta[None, :]
    U = np.maximum(0.0, (2 * w * (ap - pd) + gamma * prev["avgU"]) / kk)
    K = np.maximum(0.0, (2 * w * (am + pd) + gamma * prev["avgK"]) / kk)
    return (2 * w * p * ((U - ap + pd) - (K - am - pd))).sum(0)


def _marginal(prob, n, delta, prev, gamma, w):
    pi = prev["lam"] + (prob.ptdf[:, n][:, None] * (prev["mu"] - prev["rho"])).sum(0)
    return pi + gamma * (prev["inj"].sum(0) + delta) + _hinge_force(prob.ptdf[:, n], delta, prev, prob.fmax, gamma, w)


def generator_violation(prob, prev, P, gamma, w, prox=1.0, tol=1e-9):
    worst = 0.0
    for g in range(prob.G):
        d = P[g] - prev["P"][g]
        grad = prob.gen_mc[g] + _marginal(prob, prob.gen_node[g], d, prev, gamma, w) + prox * d
        pm = prob.gen_pmax[g]
        lo, hi = P[g] <= tol * max(1.0, pm), P[g] >= pm - tol * max(1.0, pm)
        v = np.where(lo & hi, 0.0, np.where(lo, np.maximum(0.0, -grad), np.where(hi, np.maximum(0.0, grad), np.abs(grad))))
        worst = max(worst, float(v.max()) if v.size else 0.0, float(max(0.0, -P[g].min(), (P[g] - pm).max())))
    return worst


def storage_violation(prob, prev, D, C, gamma, w, prox=1.0, tol=1e-7):
    """max over the storages of the KKT gap (0 = optimal); also checks boxes and levels"""
    worst = 0.0
    T = prob.T
    for s in range(prob.S):
        pm, em, mc = prob.sto_pmax[s], prob.sto_emax[s], prob.sto_mc[s]
        Db, Cb = prev["D"][s], prev["C"][s]
        d = (D[s] - Db) - (C[s] - Cb)
        m = _marginal(prob, prob.sto_node[s], d, prev, gamma, w)
        E = np.cumsum(C[s] - D[s])
        feas = max(0.0, -D[s].min(), -C[s].min(), (D[s] - pm).max(), (C[s] - pm).max(), -E.min(), (E - em).max())
        gD = mc + prox * (D[s] - Db) + m          # stationarity of D:  gD - eta  (= 0 | >= 0 at 0 | <= 0 at pmax)
        gC = mc + prox * (C[s] - Cb) - m          #                 C:  gC + eta
        bt = tol * max(1.0, pm)
        lo = np.full(T, -np.inf); hi = np.full(T, np.inf)
        if pm > 0:
            d0, d1 = D[s] <= bt, D[s] >= pm - bt
            hi = np.where(d1, hi, np.minimum(hi, gD))          # not at pmax: eta <= gD
            lo = np.where(d0, lo, np.maximum(lo, gD))          # not at 0:    eta >= gD
            c0, c1 = C[s] <= bt, C[s] >= pm - bt
            lo = np.where(c1, lo, np.maximum(lo, -gC))         # not at pmax: eta >= -gC
            hi = np.where(c0, hi, np.minimum(hi, -gC))         # not at 0:    eta <= -gC
        et = tol * max(1.0, em)
        flo, fhi, gap = 0.0, 0.0, 0.0                          # feasible set of eta_{t+1}
        for t in range(T - 1, -1, -1):
            at0, at1 = E[t] <= et, E[t] >= em - et
            a, b = flo, fhi
            if at0: a = -np.inf                                # eta_t <= eta_{t+1}
            if at1: b = np.inf                                 # eta_t >= eta_{t+1}
            a, b = max(a, lo[t]), min(b, hi[t])
            if a > b:                                          # empty: record the gap, continue with the midpoint
                gap = max(gap, a - b)
                a = b = 0.5 * (a + b)
            flo, fhi = a, b
        worst = max(worst, gap, feas)
    return worst


def snapshot(dev):
    """state an iteration starts from, read through the C ABI"""
    it = dev.get_iterate(); lam, mu, rho = dev.get_duals(0)
    return dict(P=it["P"], D=it["D"], C=it["C"], inj=it["injection"], flow=it["flow"], avgU=it["avgU"], avgK=it["avgK"], lam=lam, mu=mu, rho=rho)
