"""Regenerates tests/golden/*.npz from the reference's committed result traces.

Source (MIT-licensed data of rockstaedt/DecentralOPF.jl, see /root/reference/LICENSE.md):
    /root/reference/results/{TNS,big_gamma,wrong_weight}_{duals,generators,storages}.csv
written by export_results (/root/reference/src/helpers/output.jl:1-85).  Those CSVs are the only
machine-readable known answers of the reference (it has no tests).  The npz files hold exactly
the CSV numbers (float64), re-shaped to arrays:
    lam[K,T], mu[K,L,T], rho[K,L,T]   duals used BY iteration k (admm.lambdas[k] ...)
    P[K,G,T] (generator order pv, wind, coal, gas), D[K,S,T], C[K,S,T]   results OF iteration k
Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
"""
import os
import numpy as np
import pandas as pd

REF = "/root/reference/results"
OUT = os.path.dirname(os.path.abspath(__file__))
GENS = ["pv", "wind", "coal", "gas"]
STOS = ["battery"]
PARAMS = {  # which parameterisation produced which trace
    "TNS": dict(gamma=0.3, flow_weight=10.0),
    "big_gamma": dict(gamma=0.5, flow_weight=10.0),
    "wrong_weight": dict(gamma=0.3, flow_weight=0.15),
}

for name, par in PARAMS.items():
    du = pd.read_csv(f"{REF}/{name}_duals.csv")
    ge = pd.read_csv(f"{REF}/{name}_generators.csv")
    st = pd.read_csv(f"{REF}/{name}_storages.csv")
    K = int(du.iteration.max()); T = int(du.timestep.max()); L = int(du.line.max())
    lam = np.full((K, T), np.nan); mu = np.full((K, L, T), np.nan); rho = np.full((K, L, T), np.nan)
    d = du[du.dual == "lambda"]
    lam[d.iteration.values - 1, d.timestep.values - 1] = d.value.values
    for nm, arr in (("mue", mu), ("rho", rho)):
        d = du[du.dual == nm]
        arr[d.iteration.values - 1, d.line.values.astype(int) - 1, d.timestep.values - 1] = d.value.values
    P = np.full((K, len(GENS), T), np.nan)
    for gi, g in enumerate(GENS):
        d = ge[ge.generator == g]
        P[d.iteration.values - 1, gi, d.timestep.values - 1] = d.generation.values
    D = np.full((K, len(STOS), T), np.nan); Cc = np.full((K, len(STOS), T), np.nan)
    for si, s in enumerate(STOS):
        d = st[st.storage == s]
        D[d.iteration.values - 1, si, d.timestep.values - 1] = d.discharge.values
        Cc[d.iteration.values - 1, si, d.timestep.values - 1] = d.charge.values
    for a in (lam, mu, rho, P, D, Cc):
        assert not np.isnan(a).any()
    np.savez_compressed(f"{OUT}/{name}.npz", lam=lam, mu=mu, rho=rho, P=P, D=D, C=Cc,
                        gamma=par["gamma"], flow_weight=par["flow_weight"])
    print(name, "K=%d T=%d L=%d" % (K, T, L), os.path.getsize(f"{OUT}/{name}.npz"), "bytes")
