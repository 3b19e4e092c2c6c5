"""TEST-ONLY: ctypes wrapper of the sequential host emulation of the device pipeline."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libdopf_emul.so")
_dp = C.POINTER(C.c_double); _ip = C.POINTER(C.c_int)


def build():
    src = os.path.join(_HERE, "emul.cpp")
    hdrs = [os.path.join(_HERE, "../../decentralopf.jl_b200/csrc", f) for f in ("dopf_math.h", "dopf_bodies.h")]
    if os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(f) for f in [src] + hdrs):
        return _SO
    subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-o", _SO, src], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.emul_create.restype = C.c_void_p
        _lib.emul_gen_root.restype = C.c_double
        _lib.emul_exchange_buffer.restype = _dp
        _lib.emul_exchange_buffer.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_longlong)]
        _lib.emul_partition_init.argtypes = [C.c_void_p, C.c_int]
        _lib.emul_partition_finish_setup.argtypes = [C.c_void_p]
        _lib.emul_phase.argtypes = [C.c_void_p, C.c_int]
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


class EmulADMM:
    def __init__(self, prob, gamma, flow_weight=10.0, prox_weight=1.0, slack_mask_tol=1e-2, eps=1e-3, hcap=32):
        p = prob
        assert np.all(np.diff(p.gen_node) >= 0) and np.all(np.diff(p.sto_node) >= 0), "emulation expects node-sorted agents"
        self.p = p
        self.h = C.c_void_p(lib().emul_create(
            p.N, p.L, p.T, p.G, p.S, _d(p.ptdf), _d(p.fmax), _d(p.demand), _d(p.gen_mc), _d(p.gen_pmax),
            p.gen_node.ctypes.data_as(_ip), _d(p.sto_mc), _d(p.sto_pmax), _d(p.sto_emax), p.sto_node.ctypes.data_as(_ip),
            C.c_double(gamma), C.c_double(flow_weight), C.c_double(prox_weight), C.c_double(slack_mask_tol), C.c_double(eps), hcap))
        self.P = np.zeros((p.G, p.T)); self.D = np.zeros((p.S, p.T)); self.C = np.zeros((p.S, p.T)); self.E = np.zeros((p.S, p.T))
        self.inj = np.zeros((p.N, p.T)); self.flow = np.zeros((p.L, p.T)); self.avgU = np.zeros((p.L, p.T)); self.avgK = np.zeros((p.L, p.T))
        self.lam = np.zeros(p.T); self.mu = np.zeros((p.L, p.T)); self.rho = np.zeros((p.L, p.T))
        self.status = np.zeros(8, dtype=np.int32)

    def iterate(self):
        lib().emul_iterate(self.h)
        lib().emul_get(self.h, _d(self.P), _d(self.D), _d(self.C), _d(self.E), _d(self.inj), _d(self.flow), _d(self.avgU), _d(self.avgK),
                       _d(self.lam), _d(self.mu), _d(self.rho), self.status.ctypes.data_as(_ip))

    # ---- partitioned mode: mirror of dopf_set_partition / dopf_step_phase / dopf_exchange_buffer ----
    def partition_init(self, total_agents):
        lib().emul_partition_init(self.h, int(total_agents))

    def exchange_buffer(self, which):
        """numpy view (no copy) of exchange buffer 0 move maxima / 1 local injection / 2 row sums + partial flows / 3 box ranges"""
        n = C.c_longlong()
        ptr = lib().emul_exchange_buffer(self.h, which, C.byref(n))
        return np.ctypeslib.as_array(ptr, shape=(n.value,))

    def partition_finish_setup(self):
        lib().emul_partition_finish_setup(self.h)

    def phase(self, k):
        lib().emul_phase(self.h, k)

    def fetch(self):
        lib().emul_get(self.h, _d(self.P), _d(self.D), _d(self.C), _d(self.E), _d(self.inj), _d(self.flow), _d(self.avgU), _d(self.avgK),
                       _d(self.lam), _d(self.mu), _d(self.rho), self.status.ctypes.data_as(_ip))

    @property
    def iteration(self): return int(self.status[0])
    @property
    def converged(self): return bool(self.status[1])

    def __del__(self):
        try:
            lib().emul_destroy(self.h)
        except Exception:
            pass
