"""Central LP of the reference (/root/reference/src/opf_central_reference.jl:16-57) restated with scipy / HiGHS for array
problems (test infrastructure): min sum mc*P + sum mc_s*(D+C) s.t. energy balance per t, |PTDF * injection| <= f_max,
storage level recursion and boxes.  Nodal injections are explicit variables so that the constraint matrix stays sparse."""
import numpy as np
import scipy.sparse as sp
from scipy.optimize import linprog


def solve(prob):
    N, L, T, G, S = prob.N, prob.L, prob.T, prob.G, prob.S
    nP, nS, nI = G * T, S * T, N * T
    oP, oD, oC, oE, oI = 0, nP, nP + nS, nP + 2 * nS, nP + 3 * nS
    nv = oI + nI
    c = np.zeros(nv)
    c[oP:oP + nP] = np.repeat(prob.gen_mc, T)
    c[oD:oD + nS] = np.repeat(prob.sto_mc, T); c[oC:oC + nS] = np.repeat(prob.sto_mc, T)
    tt = np.arange(T)
    rows, cols, vals = [], [], []
    # I[n,t] - sum_g P - sum_s (D - C) = -demand[n,t]
    r_I = (np.arange(N)[:, None] * T + tt[None, :])
    rows += [r_I.ravel()]; cols += [oI + r_I.ravel()]; vals += [np.ones(nI)]
    gi = (np.arange(G)[:, None] * T + tt[None, :]).ravel(); gr = (prob.gen_node[:, None] * T + tt[None, :]).ravel()
    rows += [gr]; cols += [oP + gi]; vals += [-np.ones(nP)]
    si = (np.arange(S)[:, None] * T + tt[None, :]).ravel(); sr = (prob.sto_node[:, None] * T + tt[None, :]).ravel()
    rows += [sr, sr]; cols += [oD + si, oC + si]; vals += [-np.ones(nS), np.ones(nS)]
    beq = [-prob.demand.ravel()]
    nrow = nI
    # energy balance: sum_n I[n,t] = 0
    rows += [nrow + np.tile(tt, N)]; cols += [oI + r_I.ravel()]; vals += [np.ones(nI)]; beq += [np.zeros(T)]; nrow += T
    # E[s,t] - E[s,t-1] - C + D = 0
    rows += [nrow + si, nrow + si, nrow + si]; cols += [oE + si, oC + si, oD + si]; vals += [np.ones(nS), -np.ones(nS), np.ones(nS)]
    m = (np.arange(S)[:, None] * T + tt[None, 1:]).ravel()
    rows += [nrow + m]; cols += [oE + m - 1]; vals += [-np.ones(len(m))]
    beq += [np.zeros(nS)]; nrow += nS
    Aeq = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nrow, nv))
    # flows: PTDF * I[:,t] within +-fmax
    blocks = sp.kron(sp.csr_matrix(prob.ptdf), sp.identity(T, format="csr"), format="csr")      # rows l*T+t, cols n*T+t
    F = sp.hstack([sp.csr_matrix((L * T, oI)), blocks], format="csr")
    Aub = sp.vstack([F, -F], format="csr"); bub = np.concatenate([np.repeat(prob.fmax, T), np.repeat(prob.fmax, T)])
    bounds = ([(0, p) for p in np.repeat(prob.gen_pmax, T)] + [(0, p) for p in np.repeat(prob.sto_pmax, T)] * 2
              + [(0, e) for e in np.repeat(prob.sto_emax, T)] + [(None, None)] * nI)
    r = linprog(c, A_ub=Aub, b_ub=bub, A_eq=Aeq, b_eq=np.concatenate(beq), bounds=bounds, method="highs")
    if r.status != 0:
        return dict(status=r.status, message=r.message)
    x = r.x
    inj = x[oI:].reshape(N, T)
    eqm = r.eqlin.marginals
    return dict(status=0, objective=r.fun, P=x[oP:oP + nP].reshape(G, T), D=x[oD:oD + nS].reshape(S, T), C=x[oC:oC + nS].reshape(S, T),
                E=x[oE:oE + nS].reshape(S, T), injection=inj, flow=prob.ptdf @ inj, energy_price=eqm[nI:nI + T])
