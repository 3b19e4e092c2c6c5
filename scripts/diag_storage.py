"""Diagnose a GPU-vs-oracle mismatch of a storage: runs a case side by side until the first mismatch, then evaluates the TRUE
reduced objective (slacks eliminated) of both solutions for the worst storage.  The problem is strictly convex,
so the lower objective (with feasibility) tells which side is wrong.
    python scripts/diag_storage.py N L G S T wscale iters seed congest debug_flags"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
from dopf_b200.device import DeviceADMM
from oracle import oracle
N, L, G, S, T = [int(x) for x in sys.argv[1:6]]
wscale = float(sys.argv[6]); iters = int(sys.argv[7]); seed = int(sys.argv[8]); congest = float(sys.argv[9]); dbg = int(sys.argv[10])
d = pkg.cases.synthetic_arrays(N=N, L=L, G=G, S=S, T=T, seed=seed, congest_frac=congest)
prob = pkg.Problem.from_arrays(d); A = G + S
gamma, w = 0.3 / A, wscale / A
dev = DeviceADMM(prob, gamma=gamma, flow_weight=w, device=0, hinge_capacity=64, debug_flags=dbg)
ora = oracle.OracleADMM(prob, gamma, flow_weight=w)


def objective(s, Dn, Cn, prev):
    n = prob.sto_node[s]; p = prob.ptdf[:, n][:, None]                      # [L,1]
    Db, Cb = prev["D"][s], prev["C"][s]
    delta = (Dn - Db) - (Cn - Cb)                                           # [T]
    pi = prev["lam"] + (prob.ptdf[:, n][:, None] * (prev["mu"] - prev["rho"])).sum(0)
    Sbar = prev["inj"].sum(0)
    kk = 2 * w + gamma
    ap = prob.fmax[:, None] - prev["flow"]; am = prob.fmax[:, None] + prev["flow"]
    U = np.maximum(0, (2 * w * (ap - p * delta) + gamma * prev["avgU"]) / kk)
    K = np.maximum(0, (2 * w * (am + p * delta) + gamma * prev["avgK"]) / kk)
    phi = (w * (prev["flow"] + p * delta + U - prob.fmax[:, None]) ** 2 + w * (K - prev["flow"] - p * delta - prob.fmax[:, None]) ** 2
           + gamma / 2 * (U - prev["avgU"]) ** 2 + gamma / 2 * (K - prev["avgK"]) ** 2).sum(0)
    return (prob.sto_mc[s] * (Dn + Cn) + pi * (Dn - Cn) + gamma / 2 * (Sbar + delta) ** 2 + phi + 0.5 * ((Dn - Db) ** 2 + (Cn - Cb) ** 2)).sum()


for k in range(1, iters + 1):
    prev = dict(D=ora.D.copy(), C=ora.C.copy(), lam=ora.lam.copy(), mu=ora.mu.copy(), rho=ora.rho.copy(), inj=ora.inj.copy(),
                flow=ora.flow.copy(), avgU=ora.avgU.copy(), avgK=ora.avgK.copy())
    c0 = dev.status.sto_corrected; q0 = dev.status.fix_sequential
    dev.step(1); ora.iterate(0)
    it = dev.get_iterate()
    eD = np.abs(it["D"] - ora.D).max(axis=1); eC = np.abs(it["C"] - ora.C).max(axis=1); eP = np.abs(it["P"] - ora.P).max()
    print("it %d: max|dP| %.2e max|dD| %.2e max|dC| %.2e  sto corrected %d (seq %d) cold %d" % (k, eP, eD.max(), eC.max(), dev.status.sto_corrected - c0, dev.status.fix_sequential - q0, dev.status.sto_cold))
    if max(eD.max(), eC.max()) > 1e-6:
        s = int(np.argmax(np.maximum(eD, eC)))
        og = objective(s, it["D"][s], it["C"][s], prev); oo = objective(s, ora.D[s], ora.C[s], prev)
        Eg = np.cumsum(it["C"][s] - it["D"][s]); Eo = np.cumsum(ora.C[s] - ora.D[s])
        print("storage %d node %d pmax %g emax %g: objective GPU %.12g oracle %.12g  (GPU - oracle = %.3e)" % (s, prob.sto_node[s], prob.sto_pmax[s], prob.sto_emax[s], og, oo, og - oo))
        print("  GPU  level range [%.3e, %.6f], D,C range [%g,%g]" % (Eg.min(), Eg.max(), min(it["D"][s].min(), it["C"][s].min()), max(it["D"][s].max(), it["C"][s].max())))
        print("  ORA  level range [%.3e, %.6f]" % (Eo.min(), Eo.max()))
        bad = np.where(np.abs(it["C"][s] - ora.C[s]) + np.abs(it["D"][s] - ora.D[s]) > 1e-6)[0]
        print("  differing timesteps:", bad.tolist()[:40])
        for t in bad[:12]:
            print("   t=%d GPU D %.6f C %.6f E %.6f | ORA D %.6f C %.6f E %.6f | prev D %.6f C %.6f" % (t, it["D"][s][t], it["C"][s][t], Eg[t], ora.D[s][t], ora.C[s][t], Eo[t], prev["D"][s][t], prev["C"][s][t]))
        break
