"""ctypes wrapper of oracle/liboracle_b2b.so (the C restatement; test infrastructure, NOT product code).
Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_float, c_int, c_int32, c_int64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle_b2b.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            import subprocess

            subprocess.check_call(["make", "-C", _HERE, "-s"])
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.oracle_num_procs.restype = c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(c_void_p)


def _cm(a):
    """(D, N) numpy array -> Fortran-ordered float32 (Julia column-major memory)."""
    return np.asfortranarray(a, dtype=np.float32)


def num_procs() -> int:
    return lib().oracle_num_procs()


def planar_fwd(w, u, b, z, nthreads=1):
    z = _cm(z)
    D, N = z.shape
    y = np.empty((D, N), np.float32, order="F")  # fresh output per layer, like the reference
    lj = np.empty(N, np.float32)
    w = np.ascontiguousarray(w, np.float32)
    u = np.ascontiguousarray(u, np.float32)
    lib().oracle_planar_fwd_f32(_p(w), _p(u), c_float(float(np.asarray(b).reshape(-1)[0])), _p(z), _p(y), _p(lj),
                                c_int(D), c_int64(N), c_int(nthreads))
    return y, lj


def radial_fwd(alpha_raw, beta, z0, z, nthreads=1):
    z = _cm(z)
    D, N = z.shape
    y = np.empty((D, N), np.float32, order="F")
    lj = np.empty(N, np.float32)
    z0 = np.ascontiguousarray(z0, np.float32)
    lib().oracle_radial_fwd_f32(c_float(float(np.asarray(alpha_raw).reshape(-1)[0])),
                                c_float(float(np.asarray(beta).reshape(-1)[0])), _p(z0), _p(z), _p(y), _p(lj),
                                c_int(D), c_int64(N), c_int(nthreads))
    return y, lj


def rqs_fwd(W, H, Dv, x, nthreads=1):
    """W/H/Dv: (D, K1) in the reference's index order."""
    x = _cm(x)
    D, N = x.shape
    Wc, Hc, Dc = (_cm(t) for t in (W, H, Dv))
    y = np.empty((D, N), np.float32, order="F")
    lj = np.empty(N, np.float32)
    lib().oracle_rqs_fwd_f32(_p(Wc), _p(Hc), _p(Dc), c_int(W.shape[1]), _p(x), _p(y), _p(lj), c_int(D), c_int64(N),
                             c_int(nthreads))
    return y, lj


def batchnorm(b, logs, m, v, eps, x, inverse=False, nthreads=1):
    x = _cm(x)
    D, N = x.shape
    y = np.empty((D, N), np.float32, order="F")
    lj = np.empty(N, np.float32)
    arrs = [np.ascontiguousarray(t, np.float32) for t in (b, logs, m, v)]
    fn = lib().oracle_batchnorm_inv_f32 if inverse else lib().oracle_batchnorm_fwd_f32
    fn(*[_p(t) for t in arrs], c_float(float(eps)), _p(x), _p(y), _p(lj), c_int(D), c_int64(N), c_int(nthreads))
    return y, lj


def coupling_affine(idx1, idx2, W, c, x, inverse=False, nthreads=1):
    """idx1/idx2 1-based (reference convention); W (2*n1, n2)."""
    x = _cm(x)
    D, N = x.shape
    y = np.empty((D, N), np.float32, order="F")
    lj = np.empty(N, np.float32)
    i1 = np.ascontiguousarray(np.asarray(idx1) - 1, np.int32)
    i2 = np.ascontiguousarray(np.asarray(idx2) - 1, np.int32)
    Wc = _cm(W)
    cc = np.ascontiguousarray(c, np.float32)
    lib().oracle_coupling_affine_f32(_p(i1), c_int(len(i1)), _p(i2), c_int(len(i2)), _p(Wc), _p(cc),
                                     c_int(1 if inverse else 0), _p(x), _p(y), _p(lj), c_int(D), c_int64(N),
                                     c_int(nthreads))
    return y, lj


def mvnormal_diag_logpdf(mu, sigma, x, nthreads=1):
    x = _cm(x)
    D, N = x.shape
    out = np.empty(N, np.float32)
    mu = None if mu is None else np.ascontiguousarray(mu, np.float32)
    sigma = None if sigma is None else np.ascontiguousarray(sigma, np.float32)
    lib().oracle_mvnormal_diag_logpdf_f32(_p(mu), _p(sigma), _p(x), _p(out), c_int(D), c_int64(N), c_int(nthreads))
    return out


def add_inplace(y, a, nthreads=1):
    lib().oracle_add_f32(_p(y), _p(a), c_int64(y.shape[0]), c_int(nthreads))
    return y


def planar_chain_fwd(layers, x, nthreads=1):
    """with_logabsdet_jacobian(L_n ∘ … ∘ L_1, x) for planar layers [(w, u, b), ...] keeping the reference's
    pass structure: every layer allocates a fresh output and the logjac vectors are added."""
    y, lj = planar_fwd(*layers[0], x, nthreads=nthreads)
    for (w, u, b) in layers[1:]:
        y, l = planar_fwd(w, u, b, y, nthreads=nthreads)
        add_inplace(lj, l, nthreads)
    return y, lj
