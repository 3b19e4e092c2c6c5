"""TEST INFRASTRUCTURE ONLY (oracle/): ctypes harness around libdopf_oracle.so.

The oracle is the CPU restatement of the reference's ADMM iteration
(/root/reference/src/optimization/run.jl:7-16 and callees, see dopf_oracle.h).  It may be
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs only - never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libdopf_oracle.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force=False):
    """Compile the C restatement (gcc; OpenMP when libgomp is usable)."""
    srcs = [os.path.join(_HERE, f) for f in ("dopf_oracle.c", "qp_gi.c", "dopf_oracle.h", "qp_gi.h")]
    if not force and os.path.exists(_LIB_PATH) and all(
            os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return _LIB_PATH
    base = ["gcc", "-O2", "-fPIC", "-std=gnu11", "-shared", "-o", _LIB_PATH,
            srcs[0], srcs[1], "-lm"]
    for extra in (["-fopenmp"], []):
        r = subprocess.run(base[:1] + extra + base[1:], capture_output=True, text=True)
        if r.returncode == 0:
            return _LIB_PATH
    raise RuntimeError("oracle build failed:\n" + r.stderr)


class _Problem(C.Structure):
    _fields_ = [("N", C.c_int), ("L", C.c_int), ("T", C.c_int), ("G", C.c_int), ("S", C.c_int),
                ("ptdf", _dp), ("fmax", _dp), ("demand", _dp),
                ("gen_mc", _dp), ("gen_pmax", _dp), ("gen_node", _ip),
                ("sto_mc", _dp), ("sto_pmax", _dp), ("sto_emax", _dp), ("sto_node", _ip),
                ("gamma", C.c_double), ("flow_weight", C.c_double), ("prox_weight", C.c_double),
                ("slack_mask_tol", C.c_double), ("eps", C.c_double)]


class _State(C.Structure):
    _fields_ = [("iteration", C.c_int), ("converged", C.c_int),
                ("conv_lambda", C.c_int), ("conv_mue", C.c_int), ("conv_rho", C.c_int),
                ("res_lambda", C.c_double), ("res_mue", C.c_double), ("res_rho", C.c_double),
                ("total_costs", C.c_double), ("qp_kkt_worst", C.c_double),
                ("storage_outer_max", C.c_int),
                ("P", _dp), ("D", _dp), ("C", _dp), ("E", _dp), ("inj", _dp), ("flow", _dp),
                ("avgU", _dp), ("avgK", _dp), ("lam", _dp), ("mu", _dp), ("rho", _dp),
                ("lam_prev", _dp), ("mu_prev", _dp), ("rho_prev", _dp)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_iteration.restype = C.c_int
        _lib.oracle_iteration.argtypes = [C.POINTER(_Problem), C.POINTER(_State), C.c_int]
        _lib.oracle_init_state.argtypes = [C.POINTER(_Problem), C.POINTER(_State)]
        _lib.oracle_ptdf.restype = C.c_int
        _lib.oracle_ptdf.argtypes = [C.c_int, C.c_int, _ip, _ip, _dp, C.c_int, _dp]
        _lib.oracle_nodal_price.argtypes = [C.POINTER(_Problem), _dp, _dp, _dp, _dp]
        _lib.oracle_num_threads.restype = C.c_int
        _lib.qp_gi_solve.restype = C.c_int
        _lib.qp_gi_solve.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp]
        _lib.qp_kkt_residual.restype = C.c_double
        _lib.qp_kkt_residual.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp]
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _f64(a, shape=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if shape is not None:
        a = a.reshape(shape)
    return a


def num_threads():
    return int(lib().oracle_num_threads())


def set_num_threads(n):
    lib().oracle_set_num_threads(int(n))


def ptdf(N, line_from, line_to, susceptance, slack):
    """calculate_ptdf (src/helpers/ptdf.jl:1-41); 0-based indices; returns [L][N]."""
    fr = np.ascontiguousarray(line_from, dtype=np.int32)
    to = np.ascontiguousarray(line_to, dtype=np.int32)
    b = _f64(susceptance)
    out = np.zeros((len(fr), N))
    rc = lib().oracle_ptdf(N, len(fr), fr.ctypes.data_as(_ip), to.ctypes.data_as(_ip), _d(b), int(slack), _d(out))
    if rc:
        raise RuntimeError("oracle_ptdf: singular reduced susceptance matrix")
    return out


def qp_solve(G, g, Cm, b):
    G = _f64(G); g = _f64(g); Cm = _f64(Cm); b = _f64(b)
    n, m = len(g), len(b)
    x = np.zeros(n); u = np.zeros(max(m, 1))
    it = lib().qp_gi_solve(n, m, _d(G), _d(g), _d(Cm), _d(b), _d(x), _d(u))
    res = lib().qp_kkt_residual(n, m, _d(G), _d(g), _d(Cm), _d(b), _d(x), _d(u)) if it >= 0 else float("inf")
    return x, u[:m], it, res


class OracleADMM:
    """State machine equivalent to the reference's `ADMM` + `run!` (structures/admm.jl, run.jl).

    `prob` is any object/dict exposing the SoA arrays N,L,T,G,S,ptdf[L,N],fmax[L],demand[N,T],
    gen_mc,gen_pmax,gen_node,sto_mc,sto_pmax,sto_emax,sto_node.
    """

    def __init__(self, prob, gamma, flow_weight=10.0, prox_weight=1.0, slack_mask_tol=1e-2, eps=1e-3):
        g = (lambda k: prob[k]) if isinstance(prob, dict) else (lambda k: getattr(prob, k))
        self.N, self.L, self.T, self.G, self.S = (int(g(k)) for k in "NLTGS")
        N, L, T, G, S = self.N, self.L, self.T, self.G, self.S
        self._in = dict(
            ptdf=_f64(g("ptdf"), (L, N)), fmax=_f64(g("fmax"), (L,)), demand=_f64(g("demand"), (N, T)),
            gen_mc=_f64(g("gen_mc"), (G,)), gen_pmax=_f64(g("gen_pmax"), (G,)),
            gen_node=np.ascontiguousarray(g("gen_node"), dtype=np.int32).reshape(G),
            sto_mc=_f64(g("sto_mc"), (S,)), sto_pmax=_f64(g("sto_pmax"), (S,)), sto_emax=_f64(g("sto_emax"), (S,)),
            sto_node=np.ascontiguousarray(g("sto_node"), dtype=np.int32).reshape(S))
        i = self._in
        self._p = _Problem(N, L, T, G, S, _d(i["ptdf"]), _d(i["fmax"]), _d(i["demand"]),
                           _d(i["gen_mc"]), _d(i["gen_pmax"]), i["gen_node"].ctypes.data_as(_ip),
                           _d(i["sto_mc"]), _d(i["sto_pmax"]), _d(i["sto_emax"]), i["sto_node"].ctypes.data_as(_ip),
                           float(gamma), float(flow_weight), float(prox_weight), float(slack_mask_tol), float(eps))
        self.gamma = float(gamma)
        self.P = np.zeros((G, T)); self.D = np.zeros((S, T)); self.C = np.zeros((S, T)); self.E = np.zeros((S, T))
        self.inj = np.zeros((N, T)); self.flow = np.zeros((L, T))
        self.avgU = np.zeros((L, T)); self.avgK = np.zeros((L, T))
        self.lam = np.zeros(T); self.mu = np.zeros((L, T)); self.rho = np.zeros((L, T))
        self.lam_prev = np.zeros(T); self.mu_prev = np.zeros((L, T)); self.rho_prev = np.zeros((L, T))
        self._s = _State()
        self._bind()
        lib().oracle_init_state(C.byref(self._p), C.byref(self._s))

    _ARR = ("P", "D", "C", "E", "inj", "flow", "avgU", "avgK", "lam", "mu", "rho", "lam_prev", "mu_prev", "rho_prev")

    def _bind(self):
        for k in self._ARR:
            setattr(self._s, k, _d(getattr(self, k)))

    # --- reference-like read-outs ---
    @property
    def iteration(self):
        return self._s.iteration

    @property
    def converged(self):
        return bool(self._s.converged)

    @property
    def residuals(self):
        return (self._s.res_lambda, self._s.res_mue, self._s.res_rho)

    @property
    def total_costs(self):
        return self._s.total_costs

    @property
    def qp_kkt_worst(self):
        return self._s.qp_kkt_worst

    @property
    def storage_outer_max(self):
        return self._s.storage_outer_max

    def set_state(self, **arrs):
        """Overwrite parts of the state (arrays named as the attributes; `iteration` allowed)."""
        for k, v in arrs.items():
            if k == "iteration":
                self._s.iteration = int(v)
            else:
                getattr(self, k)[...] = v

    def iterate(self, mode=0):
        """calculate_iteration!(admm) (run.jl:7-16)."""
        rc = lib().oracle_iteration(C.byref(self._p), C.byref(self._s), int(mode))
        if rc:
            raise RuntimeError(f"oracle_iteration failed rc={rc}")

    def run(self, max_iterations=100000, mode=0):
        """run!(admm) (run.jl:1-5) with an iteration cap."""
        n = 0
        while not self.converged and n < max_iterations:
            self.iterate(mode)
            n += 1
        return n

    def nodal_price(self, which="prev"):
        """get_nodal_price(admm.iteration) (network_elements.jl:16-25): uses the duals of the
        last executed iteration (`prev`) when called right after convergence as the driver does."""
        lam, mu, rho = (self.lam_prev, self.mu_prev, self.rho_prev) if which == "prev" else (self.lam, self.mu, self.rho)
        out = np.zeros((self.N, self.T))
        lib().oracle_nodal_price(C.byref(self._p), _d(lam), _d(mu), _d(rho), _d(out))
        return out
