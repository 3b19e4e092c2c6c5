"""The device math (csrc/dopf_math.h, dopf_bodies.h) compiled for the host with a lane group of
width 1 and run sequentially (tests/host_emul) against the oracle - lets the per-agent
algorithms and the predict/verify/correct pipeline be checked without a GPU."""
import ctypes as C

import numpy as np
import pytest

from tests.conftest import GOLDEN, load_golden

dp = C.POINTER(C.c_double); ip = C.POINTER(C.c_int)
d_ = lambda a: a.ctypes.data_as(dp)


@pytest.fixture(scope="module")
def emul():
    from tests.host_emul import emul as e
    e.build()
    return e


def _storage_qp(oracle_mod, T, mc, pmax, emax, prox, Db, Cb, g0, s1):
    n = 2 * T; G = np.zeros((n, n)); g = np.zeros(n)
    for t in range(T):
        a = s1[t]; dbar = Db[t] - Cb[t]
        G[t, t] = G[T + t, T + t] = prox + a; G[t, T + t] = G[T + t, t] = -a
        net0 = g0[t] - a * dbar
        g[t] = mc + net0 - prox * Db[t]; g[T + t] = mc - net0 - prox * Cb[t]
    rows, b = [], []
    for t in range(T):
        for i, s, bb in ((t, 1, 0), (t, -1, -pmax), (T + t, 1, 0), (T + t, -1, -pmax)):
            r = np.zeros(n); r[i] = s; rows.append(r); b.append(bb)
    for t in range(T):
        r = np.zeros(n); r[T:T + t + 1] = 1; r[:t + 1] = -1
        rows.append(r); b.append(0); rows.append(-r); b.append(-emax)
    x, u, it, res = oracle_mod.qp_solve(G, g, np.array(rows), np.array(b, dtype=float))
    assert it >= 0 and res < 1e-6
    return x[:T], x[T:]


@pytest.mark.parametrize("fn", ["emul_storage_solve", "emul_storage_solve_seq"])
def test_storage_funnel_equals_dense_qp(emul, oracle_mod, fn):
    lib = emul.lib()
    rng = np.random.default_rng(1)
    for trial in range(150):
        T = int(rng.choice([1, 2, 3, 5, 8, 24, 48]))
        mc = float(rng.choice([0.0, 1.0, 3.0])); pmax = float(rng.integers(5, 50)); emax = pmax * float(rng.choice([0.5, 1, 2, 4])); prox = float(rng.choice([1.0, 0.3, 5.0]))
        mode = trial % 4
        Db = rng.uniform(0, pmax, T) * (rng.random(T) < 0.5); Cb = rng.uniform(0, pmax, T) * (rng.random(T) < 0.5)
        amp = [1.0, 10.0, 40.0, 100.0][mode]
        g0 = rng.normal(0, amp, T) + amp * np.sin(np.arange(T) / 3.0); s1 = rng.uniform(0, [0.01, 0.3, 3.0, 30.0][mode], T)
        D = np.zeros(T); Cc = np.zeros(T); eta = np.zeros(T); st = np.zeros(4, dtype=np.int32)
        getattr(lib, fn)(T, C.c_double(mc), C.c_double(pmax), C.c_double(emax), C.c_double(prox), d_(Db), d_(Cb), d_(g0), d_(s1),
                         0, None, None, None, d_(D), d_(Cc), d_(eta), st.ctypes.data_as(ip))
        Dq, Cq = _storage_qp(oracle_mod, T, mc, pmax, emax, prox, Db, Cb, g0, s1)
        assert max(np.abs(D - Dq).max(), np.abs(Cc - Cq).max()) < 1e-8, (trial, T)


def test_storage_warm_start_is_exact_or_falls_back(emul, oracle_mod):
    """drifting prices over 12 'iterations': whenever the warm path verifies it must equal the QP"""
    lib = emul.lib()
    rng = np.random.default_rng(5)
    nwarm = 0
    for trial in range(25):
        T = int(rng.choice([6, 24, 48])); pmax = float(rng.integers(5, 51)); emax = pmax * float(rng.integers(2, 5))
        Db = np.zeros(T); Cb = np.zeros(T); eta = np.zeros(T); Eprev = np.zeros(T)
        base = -30 + 20 * np.sin(np.arange(T) * 2 * np.pi / 24) + rng.normal(0, 3, T); s1 = np.full(T, 3e-4)
        for rep in range(12):
            g0 = base + rng.normal(0, 0.3 if rep > 3 else 3.0, T)
            D = np.zeros(T); Cc = np.zeros(T); st = np.zeros(4, dtype=np.int32)
            w = lib.emul_storage_solve_warm(T, C.c_double(1.0), C.c_double(pmax), C.c_double(emax), C.c_double(1.0), d_(Db), d_(Cb), d_(g0), d_(s1),
                                            d_(D), d_(Cc), d_(eta), d_(Eprev), st.ctypes.data_as(ip))
            Dq, Cq = _storage_qp(oracle_mod, T, 1.0, pmax, emax, 1.0, Db, Cb, g0, s1)
            assert max(np.abs(D - Dq).max(), np.abs(Cc - Cq).max()) < 1e-8, (trial, rep, w)
            nwarm += w
            Db, Cb = D.copy(), Cc.copy(); Eprev = np.cumsum(Cc - D)
    assert nwarm > 50      # the warm path is actually exercised


def test_generator_root_with_hinges(emul):
    lib = emul.lib()
    rng = np.random.default_rng(3)
    for trial in range(300):
        n = int(rng.integers(0, 9))
        bp = rng.normal(0, 5, n); s = rng.uniform(0.1, 30, n); sg = s * rng.choice([-1.0, 1.0], n)
        c = float(rng.normal(0, 20)); lo, hi = -float(rng.uniform(0, 30)), float(rng.uniform(0, 30))
        # slope a must exceed the anchored hinge slopes (they are already part of it)
        anchored = (np.sign(sg) * bp) < 0
        a = 1.0 + s[anchored].sum() + float(rng.uniform(0, 2))
        x = lib.emul_gen_root(C.c_double(c), C.c_double(a), n, d_(bp), d_(sg), C.c_double(lo), C.c_double(hi))

        def f(z):
            v = c + a * z
            for b, w in zip(bp, sg):
                dirn = 1.0 if w > 0 else -1.0
                e = dirn * (z - b)
                if dirn * b < 0:
                    v -= abs(w) * (z - b) if e < 0 else 0.0
                else:
                    v += abs(w) * (z - b) if e > 0 else 0.0
            return v
        assert lo - 1e-12 <= x <= hi + 1e-12
        if lo + 1e-9 < x < hi - 1e-9:
            assert abs(f(x)) < 1e-9 * (1 + abs(c))
        elif x <= lo + 1e-9:
            assert f(lo) >= -1e-9
        else:
            assert f(hi) <= 1e-9


@pytest.mark.parametrize("fn", ["emul_storage_solve", "emul_storage_solve_seq"])
def test_storage_with_hinges_minimises_the_true_objective(emul, fn):
    """storage solve with explicit slack hinges (the correction pass): the result must minimise
    sum_t mc(D+C) + prox/2((D-Db)^2+(C-Cb)^2) + phi_t(delta_t) over the boxes and level bounds, where
    phi_t' = g0 + s1*delta + sum of hinge terms; checked against SLSQP on the C1 objective."""
    from scipy.optimize import minimize
    lib = emul.lib()
    rng = np.random.default_rng(11)
    hcap = 6
    for trial in range(60):
        T = int(rng.choice([1, 2, 3, 5]))
        mc = float(rng.choice([0.0, 1.0])); pmax = float(rng.integers(5, 30)); emax = pmax * float(rng.choice([0.5, 1, 2])); prox = float(rng.choice([1.0, 0.5]))
        Db = rng.uniform(0, pmax, T) * (rng.random(T) < 0.5); Cb = rng.uniform(0, pmax, T) * (rng.random(T) < 0.5)
        g0 = rng.normal(0, 15, T)
        hcnt = rng.integers(0, hcap + 1, T).astype(np.int32)
        hbp = rng.normal(0, 0.4 * pmax, (T, hcap)); hs = rng.uniform(0.1, 5, (T, hcap)); hsg = hs * rng.choice([-1.0, 1.0], (T, hcap))
        s1 = np.zeros(T)
        for t in range(T):
            k = hcnt[t]
            anchored = (np.sign(hsg[t, :k]) * hbp[t, :k]) < 0
            s1[t] = float(rng.uniform(0.0, 1.0)) + hs[t, :k][anchored].sum()
        D = np.zeros(T); Cc = np.zeros(T); eta = np.zeros(T); st = np.zeros(4, dtype=np.int32)
        getattr(lib, fn)(T, C.c_double(mc), C.c_double(pmax), C.c_double(emax), C.c_double(prox), d_(Db), d_(Cb), d_(g0), d_(s1),
                         hcap, hcnt.ctypes.data_as(ip), d_(np.ascontiguousarray(hbp)), d_(np.ascontiguousarray(hsg)),
                         d_(D), d_(Cc), d_(eta), st.ctypes.data_as(ip))

        def phi(t, z):
            v = g0[t] * z + 0.5 * s1[t] * z * z
            for b, w in zip(hbp[t, :hcnt[t]], hsg[t, :hcnt[t]]):
                dirn = 1.0 if w > 0 else -1.0
                e = dirn * (z - b)
                if dirn * b < 0:
                    v -= 0.5 * abs(w) * (z - b) ** 2 if e < 0 else 0.0
                else:
                    v += 0.5 * abs(w) * (z - b) ** 2 if e > 0 else 0.0
            return v

        def obj(x):
            d, c = x[:T], x[T:]
            return sum(mc * (d[t] + c[t]) + 0.5 * prox * ((d[t] - Db[t]) ** 2 + (c[t] - Cb[t]) ** 2) + phi(t, (d[t] - Db[t]) - (c[t] - Cb[t])) for t in range(T))
        lev = lambda x: np.cumsum(x[T:] - x[:T])
        cons = [{"type": "ineq", "fun": lambda x: lev(x)}, {"type": "ineq", "fun": lambda x: emax - lev(x)}]
        x0 = np.concatenate([D, Cc])
        assert lev(x0).min() >= -1e-9 and lev(x0).max() <= emax + 1e-9 and x0.min() >= -1e-12 and x0.max() <= pmax + 1e-12
        best = obj(x0)
        for start in (x0, np.zeros(2 * T), np.concatenate([Db, Cb]) * 0.5):
            r = minimize(obj, start, method="SLSQP", bounds=[(0, pmax)] * (2 * T), constraints=cons, options={"ftol": 1e-14, "maxiter": 500})
            feas = lev(r.x).min() >= -1e-7 and lev(r.x).max() <= emax + 1e-7
            if r.success and feas:
                assert best <= r.fun + 1e-6 * (1 + abs(r.fun)), (trial, T, best, r.fun)


@pytest.mark.parametrize("name", GOLDEN)
def test_pipeline_emulation_reproduces_reference_trace(emul, oracle_mod, three_node_sorted, name):
    prob, order = three_node_sorted
    g = load_golden(name)
    e = emul.EmulADMM(prob, float(g["gamma"]), flow_weight=float(g["flow_weight"]))
    o = oracle_mod.OracleADMM(prob, float(g["gamma"]), flow_weight=float(g["flow_weight"]))
    worst_g = worst_o = 0.0
    for k in range(min(g["P"].shape[0], 500)):
        e.iterate(); o.iterate(0)
        if e.converged:
            assert name == "TNS" and k + 1 == 476
            break
        worst_g = max(worst_g, np.abs(e.P - g["P"][k][order]).max(), np.abs(e.D - g["D"][k]).max(), np.abs(e.C - g["C"][k]).max())
        worst_o = max(worst_o, np.abs(e.P - o.P).max(), np.abs(e.mu - o.mu).max(), np.abs(e.rho - o.rho).max(), np.abs(e.avgU - o.avgU).max())
    assert worst_g < 2e-5 and worst_o < 1e-9
    assert e.status[6] == 0


@pytest.mark.parametrize("dims,gamma,w,iters", [((12, 18, 30, 8, 6), None, None, 40), ((12, 18, 30, 8, 6), 0.02, 10.0, 25),
                                                 ((40, 60, 200, 40, 24), None, None, 15)])
def test_pipeline_emulation_equals_oracle_on_synthetic(pkg, emul, oracle_mod, dims, gamma, w, iters):
    N, L, G, S, T = dims
    d = pkg.cases.synthetic_arrays(N=N, L=L, G=G, S=S, T=T, seed=3)
    prob = pkg.Problem.from_arrays(d)
    A = G + S
    gamma = gamma or 0.3 / A; w = w or 1.0 / A
    o = oracle_mod.OracleADMM(prob, gamma, flow_weight=w); e = emul.EmulADMM(prob, gamma, flow_weight=w, hcap=64)
    for k in range(iters):
        o.iterate(0); e.iterate()
        flips = ((o.mu == 0) != (e.mu == 0)).sum() + ((o.rho == 0) != (e.rho == 0)).sum()
        assert flips == 0, f"slack-mask flip at iteration {k + 1}"       # reported separately from drift
        for a, b in ((o.P, e.P), (o.D, e.D), (o.C, e.C), (o.lam, e.lam), (o.mu, e.mu), (o.rho, e.rho), (o.avgU, e.avgU), (o.avgK, e.avgK)):
            assert np.abs(a - b).max() <= 1e-6 * max(1.0, np.abs(a).max())
    assert e.status[6] == 0 and e.status[2] > 0       # the correction pass was exercised


def test_clip_table_closed_form_equals_evaluated_table(emul):
    """The per-timestep clip table of the warp storage solver: the closed form (sto_clip_table, what the device builds) against
    the table built from evaluations of D, C at the breakpoints and piece midpoints, and both against the defining equation
    nu = g0 - eta + s1 * delta(nu) at random multipliers - including degenerate steps (pmax = 0, coinciding breakpoints)."""
    import ctypes as C
    lib = emul.lib()
    dbl = C.c_double
    lib.emul_clip_table.argtypes = [dbl] * 7 + [C.c_int, C.POINTER(dbl)]
    lib.emul_clip_eval.argtypes = [C.POINTER(dbl)] + [dbl] * 6 + [C.POINTER(dbl)]
    rng = np.random.default_rng(7)
    worst_tab = worst_eq = 0.0
    for trial in range(4000):
        pmax = float(rng.choice([0.0, 1e-9, 5.0, 37.0, 50.0]))
        mc = float(rng.choice([0.0, 1.0, 3.5])); prox = float(rng.choice([1.0, 0.25, 4.0]))
        Db = float(rng.choice([0.0, pmax, rng.uniform(0, pmax)])); Cb = float(rng.choice([0.0, pmax, rng.uniform(0, pmax), Db]))
        g0 = float(rng.normal(0, 30)); s1 = float(rng.choice([3e-7, 1e-3, 0.3, 2.0]))
        a = np.zeros(19); b = np.zeros(19)
        lib.emul_clip_table(Db, Cb, g0, s1, mc, pmax, prox, 0, a.ctypes.data_as(C.POINTER(dbl)))
        lib.emul_clip_table(Db, Cb, g0, s1, mc, pmax, prox, 1, b.ctypes.data_as(C.POINTER(dbl)))
        scale = 1.0 + abs(g0) + prox * pmax + mc
        assert np.all(np.diff(a[:4]) <= 1e-12 * scale), (trial, a[:4])            # thresholds non-increasing
        worst_tab = max(worst_tab, np.abs(a[:4] - b[:4]).max() / scale)
        etas = np.concatenate([rng.normal(g0, 3 * scale, 6), a[:4], a[:4] + 1e-9 * scale, a[:4] - 1e-9 * scale])
        for eta in etas:
            oa = np.zeros(4); ob = np.zeros(4)
            lib.emul_clip_eval(a.ctypes.data_as(C.POINTER(dbl)), Db, Cb, mc, pmax, prox, float(eta), oa.ctypes.data_as(C.POINTER(dbl)))
            lib.emul_clip_eval(b.ctypes.data_as(C.POINTER(dbl)), Db, Cb, mc, pmax, prox, float(eta), ob.ctypes.data_as(C.POINTER(dbl)))
            D, Cc, dy, nu = oa
            # the defining equation of the hinge-free step
            worst_eq = max(worst_eq, abs(nu - (g0 - eta + s1 * ((D - Db) - (Cc - Cb)))) / (scale + abs(eta)))
            assert abs(D - ob[0]) <= 1e-10 * (1 + pmax) and abs(Cc - ob[1]) <= 1e-10 * (1 + pmax), (trial, eta, oa, ob)
    assert worst_tab < 1e-13 and worst_eq < 1e-13, (worst_tab, worst_eq)
