"""PTDF construction (setup step, host side).

Same construction as /root/reference/src/helpers/ptdf.jl:1-41: incidence A (L x N, +1 at
`from`, -1 at `to`), B = diag(susceptance), Bl = B*A, Bn = A'*B*A, inverse of Bn without the
slack row/column embedded in zeros, PTDF = Bl * B_inv (slack column = 0).  The first node with
slack == true is the slack (ptdf.jl:12-16).  A linear solve replaces the dense `inv`.
"""
import numpy as np


def ptdf_from_arrays(N, line_from, line_to, susceptance, slack):
    L = len(line_from)
    A = np.zeros((L, N))
    A[np.arange(L), np.asarray(line_from)] = 1.0
    A[np.arange(L), np.asarray(line_to)] = -1.0
    b = np.asarray(susceptance, dtype=np.float64)
    Bl = b[:, None] * A
    Bn = A.T @ Bl
    keep = np.array([n for n in range(N) if n != slack])
    out = np.zeros((L, N))
    # PTDF[:, keep] = Bl[:, keep] * inv(Bn[keep, keep])  <=>  solve Bn_kk' X' = Bl_k'
    out[:, keep] = np.linalg.solve(Bn[np.ix_(keep, keep)].T, Bl[:, keep].T).T
    return out


def ptdf_device(N, line_from, line_to, susceptance, slack, device=-1):
    """the same PTDF built on the GPU (dopf_calculate_ptdf: Cholesky + triangular solves of the slack-reduced susceptance
    matrix); raises without a CUDA device"""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    fr = np.ascontiguousarray(line_from, dtype=np.int32); to = np.ascontiguousarray(line_to, dtype=np.int32)
    b = np.ascontiguousarray(susceptance, dtype=np.float64)
    out = np.empty((len(fr), N))
    rc = lib.dopf_calculate_ptdf(N, len(fr), fr.ctypes.data_as(C.c_void_p), to.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
                                 int(slack), int(device), out.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError(f"dopf_calculate_ptdf rc={rc}: {lib.dopf_ptdf_last_error().decode()}")
    return out


def calculate_ptdf(nodes, lines):
    """calculate_ptdf(nodes, lines) -> [L, N] float64 (ptdf.jl:1-41)."""
    idx = {id(n): i for i, n in enumerate(nodes)}
    slack = next(i for i, n in enumerate(nodes) if n.slack)
    fr = [idx[id(l.from_)] for l in lines]
    to = [idx[id(l.to)] for l in lines]
    return ptdf_from_arrays(len(nodes), fr, to, [l.susceptance for l in lines], slack)
