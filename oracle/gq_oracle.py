"""ctypes/numpy front-end of the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package (gradient-quantization_b200/)
never does: it fails loudly when its CUDA library is missing.

The arithmetic lives in gq_oracle.c (each function cites the reference lines it
restates).  This file only marshals numpy arrays and composes the per-tensor
codecs into the quantizers' record/apply flow
(reference quantizers/ps_quantizer.py:27-65, quantizers/ring_quantizer.py:25-49).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgq_oracle.so")


def build(force=False):
    """Compile gq_oracle.c -> libgq_oracle.so (gcc, a second or two)."""
    src = os.path.join(_HERE, "gq_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


_lib = None
_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.gqo_psc_compress.restype = ctypes.c_int
        _lib.gqo_hsq_roundtrip.restype = ctypes.c_int
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_f32p)


def _i64(x):
    return ctypes.c_int64(int(x))


# ----------------------------------------------------------------------------
# dim rule shared by HSQ and QSGD
# (nearest_neighbor_compressor.py:23-29, qsgd_compressor.py:16-22)
def chunk_dim(size, c_dim):
    if c_dim == 0 or size < c_dim:
        return size
    dim = c_dim
    for _ in range(10):
        if size % dim != 0:
            dim = dim // 2 * 3
    return dim


def fvecs_read(path):
    """utils/vecs_io.py:5-12: int32 d, then d fp32, per row."""
    a = np.fromfile(path, dtype="int32")
    d = a[0]
    return a.reshape(-1, d + 1)[:, 1:].copy().view("float32")


def normalize(vecs):
    """utils/vec_np.py:4-10 (same numpy calls: the codebook fed to the kernels is
    produced by identical host code, so it is bit-identical to the reference's)."""
    norms = np.linalg.norm(vecs, axis=1)
    nm = norms[:, np.newaxis]
    return norms, np.divide(vecs, nm, out=np.zeros_like(vecs), where=nm != 0)


# ----------------------------------------------------------------------------
def hsq_search(v, cb):
    """a2. v: [nchunks, d] fp32, cb: [K, d] fp32 -> (codes int32[nchunks], u fp32[nchunks])."""
    v, vp = _f(v)
    nchunks, d = v.shape
    cbT, cbTp = _f(np.ascontiguousarray(cb.T))
    K = cb.shape[0]
    codes = np.empty(nchunks, dtype=np.int32)
    u = np.empty(nchunks, dtype=np.float32)
    lib().gqo_hsq_search(vp, _i64(nchunks), d, cbTp, K,
                         codes.ctypes.data_as(_i32p), u.ctypes.data_as(_f32p))
    return codes, u


def hsq_scores(v, cb):
    v, vp = _f(v)
    nchunks, d = v.shape
    cbT, cbTp = _f(np.ascontiguousarray(cb.T))
    K = cb.shape[0]
    p = np.empty((nchunks, K), dtype=np.float32)
    lib().gqo_hsq_scores(vp, _i64(nchunks), d, cbTp, K, p.ctypes.data_as(_f32p))
    return p


def psc_compress(u, n_bit, random=True, r=None):
    """a4 -> (lb, ub, l int32, consumed_uniforms: bool)."""
    u, up = _f(u)
    n = u.size
    l = np.empty(n, dtype=np.int32)
    lb = ctypes.c_float()
    ub = ctypes.c_float()
    if random:
        r, rp = _f(r)
        assert r.size >= n
    else:
        rp = None
    used = lib().gqo_psc_compress(up, _i64(n), n_bit, int(bool(random)), rp,
                                  ctypes.byref(lb), ctypes.byref(ub), l.ctypes.data_as(_i32p))
    return np.float32(lb.value), np.float32(ub.value), l, bool(used)


def psc_decompress(lb, ub, l, n_bit):
    """a5."""
    l = np.ascontiguousarray(l, dtype=np.int32)
    out = np.empty(l.size, dtype=np.float32)
    lib().gqo_psc_decompress(l.ctypes.data_as(_i32p), _i64(l.size), n_bit,
                             ctypes.c_float(float(lb)), ctypes.c_float(float(ub)),
                             out.ctypes.data_as(_f32p))
    return out


def hsq_decode(codes, norms, cb):
    """a3 -> fp32 [nchunks*d]."""
    codes = np.ascontiguousarray(codes, dtype=np.int32)
    norms, np_ = _f(norms)
    cb, cbp = _f(cb)
    d = cb.shape[1]
    out = np.empty(codes.size * d, dtype=np.float32)
    lib().gqo_hsq_decode(codes.ctypes.data_as(_i32p), np_, _i64(codes.size), d, cbp,
                         out.ctypes.data_as(_f32p))
    return out


def qsgd_compress(v, dim, n_bit, random=True, r=None):
    """a8 -> (norm fp32[M], signs uint8[N], l int32[N])."""
    v, vp = _f(np.asarray(v).reshape(-1))
    n = v.size
    m = n // dim
    norm = np.empty(m, dtype=np.float32)
    signs = np.empty(n, dtype=np.uint8)
    l = np.empty(n, dtype=np.int32)
    if random:
        r, rp = _f(np.asarray(r).reshape(-1))
        assert r.size >= n
    else:
        rp = None
    lib().gqo_qsgd_compress(vp, _i64(m), dim, n_bit, int(bool(random)), rp,
                            norm.ctypes.data_as(_f32p), signs.ctypes.data_as(_u8p),
                            l.ctypes.data_as(_i32p))
    return norm, signs, l


def qsgd_decompress(norm, signs, l, dim, n_bit):
    norm, np_ = _f(norm)
    signs = np.ascontiguousarray(signs, dtype=np.uint8)
    l = np.ascontiguousarray(l, dtype=np.int32)
    out = np.empty(l.size, dtype=np.float32)
    lib().gqo_qsgd_decompress(np_, signs.ctypes.data_as(_u8p), l.ctypes.data_as(_i32p),
                              _i64(norm.size), dim, n_bit, out.ctypes.data_as(_f32p))
    return out


def sign(v):
    v, vp = _f(np.asarray(v).reshape(-1))
    out = np.empty(v.size, dtype=np.float32)
    lib().gqo_sign(vp, _i64(v.size), out.ctypes.data_as(_f32p))
    return out


def topk(v, k):
    v, vp = _f(np.asarray(v).reshape(-1))
    out = np.empty(v.size, dtype=np.float32)
    lib().gqo_topk(vp, _i64(v.size), _i64(k), out.ctypes.data_as(_f32p))
    return out


def pvc_search(v, dagger, r):
    """a7 (intended semantics, parity unpinned). dagger = pinv(C^T): [K, d]."""
    v, vp = _f(v)
    nchunks, d = v.shape
    dT, dTp = _f(np.ascontiguousarray(dagger.T))
    K = dagger.shape[0]
    r, rp = _f(r)
    codes = np.empty(nchunks, dtype=np.int32)
    u = np.empty(nchunks, dtype=np.float32)
    lib().gqo_pvc_search(vp, _i64(nchunks), d, dTp, K, rp,
                         codes.ctypes.data_as(_i32p), u.ctypes.data_as(_f32p))
    return codes, u


def ps_mean(users):
    """a13 apply(): users [U, n] -> mean over users (sum in user order, then / U)."""
    users, up = _f(users)
    U, n = users.shape
    out = np.empty(n, dtype=np.float32)
    lib().gqo_ps_mean(up, U, _i64(n), out.ctypes.data_as(_f32p))
    return out


# ----------------------------------------------------------------------------
# Per-tensor codecs with the reference's signature containers, numpy flavoured.
class UniformStream:
    """Replays a pre-drawn uniform stream in reference call order
    (user-major, tensor-minor; one rand(N/d) per HSQ tensor with
    lb != ub, one rand(M, dim) per QSGD tensor)."""

    def __init__(self, draws):
        self.draws = np.ascontiguousarray(draws, dtype=np.float32).reshape(-1)
        self.pos = 0

    def peek(self, n):
        assert self.pos + n <= self.draws.size, "uniform stream exhausted"
        return self.draws[self.pos:self.pos + n]

    def advance(self, n):
        self.pos += n


class HSQ:
    """NearestNeighborCompressor (a1-a3) over the oracle primitives."""

    def __init__(self, size, shape, cb, n_bit, random=True):
        self.size, self.shape = size, tuple(shape)
        self.cb = np.ascontiguousarray(cb, dtype=np.float32)
        self.K, self.dim = self.cb.shape
        assert size % self.dim == 0
        self.n_bit, self.random = n_bit, random

    def compress(self, vec, stream=None):
        v = np.asarray(vec, dtype=np.float32).reshape(-1, self.dim)
        codes, u = hsq_search(v, self.cb)
        if self.n_bit == 32:
            return [u, codes]
        r = stream.peek(u.size) if self.random else None
        lb, ub, l, used = psc_compress(u, self.n_bit, self.random, r)
        if used:
            stream.advance(u.size)
        return [(lb, ub, l), codes]

    def decompress(self, sig):
        norms, codes = sig
        if self.n_bit != 32:
            norms = psc_decompress(norms[0], norms[1], norms[2], self.n_bit)
        return hsq_decode(codes, norms, self.cb).reshape(self.shape)


class QSGD:
    def __init__(self, size, shape, c_dim, n_bit, random=True):
        self.size, self.shape = size, tuple(shape)
        self.dim = chunk_dim(size, c_dim)
        assert self.dim != 0 and size % self.dim == 0
        self.n_bit, self.random = n_bit, random

    def compress(self, vec, stream=None):
        r = stream.peek(self.size) if self.random else None
        out = qsgd_compress(vec, self.dim, self.n_bit, self.random, r)
        if self.random:
            stream.advance(self.size)
        return list(out)

    def decompress(self, sig):
        return qsgd_decompress(sig[0], sig[1], sig[2], self.dim, self.n_bit).reshape(self.shape)


class Sign:
    def __init__(self, size, shape):
        self.shape = tuple(shape)

    def compress(self, vec, stream=None):
        return sign(vec)

    def decompress(self, sig):
        return sig.reshape(self.shape)


def sign_pack_t5(sig):
    """Base-3 wire of a ternary sign vector (include/gqb200.h gq_sign_encode_t5): five elements per byte,
    byte = t0 + 3 t1 + 9 t2 + 27 t3 + 81 t4, digit 0 -> 0, 1 -> +1, 2 -> -1; padded with zeros to a
    multiple of 20 elements (whole 32-bit words)."""
    v = np.asarray(sig, np.float32).reshape(-1)
    n = v.size
    t = np.zeros((n + 19) // 20 * 20, np.uint8)
    t[:n] = np.where(v > 0, 1, np.where(v < 0, 2, 0))
    w = np.array([1, 3, 9, 27, 81], np.uint16)
    return (t.reshape(-1, 5).astype(np.uint16) * w).sum(1).astype(np.uint8)


def sign_unpack_t5(packed, n):
    b = np.asarray(packed, np.uint8).astype(np.int32)
    digits = np.stack([(b // d) % 3 for d in (1, 3, 9, 27, 81)], axis=1).reshape(-1)[:n]
    return np.where(digits == 1, 1.0, np.where(digits == 2, -1.0, 0.0)).astype(np.float32)


class TopK:
    def __init__(self, size, shape, cr):
        self.shape, self.k = tuple(shape), size // cr

    def compress(self, vec, stream=None):
        return topk(vec, self.k)

    def decompress(self, sig):
        return sig.reshape(self.shape)


class Identity:
    def compress(self, vec, stream=None):
        return np.array(vec, dtype=np.float32, copy=True)

    def decompress(self, sig):
        return sig


class PVC:
    """ProbabilisticVectorCompressor (a7).  Pinned against the reference's own arithmetic for every
    line that executes (tests/golden/make_golden.py:pvc_case); the one line that does not --
    `argmin` on a bool tensor, :57-58 -- is read as "first index whose cumulative probability
    reaches r - 1e-5" (what argmin + 1 gave on the ByteTensors of the PyTorch it was written for)."""

    def __init__(self, size, shape, cb, n_bit, random=True):
        self.size, self.shape = size, tuple(shape)
        self.cb = np.ascontiguousarray(cb, dtype=np.float32)
        self.K, self.dim = self.cb.shape
        self.dagger = np.linalg.pinv(self.cb.T).astype(np.float32)
        self.n_bit, self.random = n_bit, random

    def compress(self, vec, stream):
        v = np.asarray(vec, dtype=np.float32).reshape(-1, self.dim)
        r = stream.peek(v.shape[0])
        stream.advance(v.shape[0])
        codes, u = pvc_search(v, self.dagger, r)
        if self.n_bit == 32:
            return [u, codes]
        r2 = stream.peek(u.size) if self.random else None
        lb, ub, l, used = psc_compress(u, self.n_bit, self.random, r2)
        if used:
            stream.advance(u.size)
        return [(lb, ub, l), codes]

    def decompress(self, sig):
        norms, codes = sig
        if self.n_bit != 32:
            norms = psc_decompress(norms[0], norms[1], norms[2], self.n_bit)
        return hsq_decode(codes, norms, self.cb).reshape(self.shape)


class Residual:
    """ResidualCompressor (a6): stage 1 HSQ, stage 2 PVC on the residual."""

    def __init__(self, size, shape, cb, n_bit, random=True):
        self.shape = tuple(shape)
        self.stages = [HSQ(size, shape, cb, n_bit, random), PVC(size, shape, cb, n_bit, random)]

    def compress(self, vec, stream=None):
        res = np.array(vec, dtype=np.float32, copy=True).reshape(self.shape)
        sigs = []
        for c in self.stages:
            sig = c.compress(res, stream)
            res = res - c.decompress(sig)
            sigs.append(sig)
        return sigs

    def decompress(self, sigs):
        out = self.stages[0].decompress(sigs[0])
        for c, s in zip(self.stages[1:], sigs[1:]):
            out = out + c.decompress(s)
        return out


def ps_scale(epoch, scale="exp"):
    """ps_quantizer.py:28-31."""
    import math
    return (2 / (math.exp(-epoch) + 1) - 1) if scale == "exp" else float(scale)


def ps_step(codecs, user_grads, stream=None, ef_errors=None, scale=0.0, two_phase=False, server_errors=None):
    """ps_quantizer.py:27-65: user_grads[u][i] -> the gradients apply() leaves in param.grad.
    ef_errors[u][i] (updated in place) enables error feedback (:34-39); two_phase compresses the
    average once more (:52-61), with server_errors[i] (updated in place) when error feedback is on."""
    U = len(user_grads)
    out = []
    dec = [[None] * len(codecs) for _ in range(U)]
    for u in range(U):
        for i, c in enumerate(codecs):
            g = np.asarray(user_grads[u][i], dtype=np.float32)
            if ef_errors is not None:
                g = g + np.float32(scale) * ef_errors[u][i]
            d = c.decompress(c.compress(g, stream))
            if ef_errors is not None:
                ef_errors[u][i] = g - d
            dec[u][i] = d
    for i, c in enumerate(codecs):
        stack = np.stack([dec[u][i].reshape(-1) for u in range(U)])
        g = ps_mean(stack).reshape(dec[0][i].shape)
        if two_phase:
            if server_errors is not None:
                g = g + server_errors[i]
                d = c.decompress(c.compress(g, stream))
                server_errors[i] = g - d
                g = d
            else:
                g = c.decompress(c.compress(g, stream))
        out.append(g)
    return out


def ring_step(codecs, user_grads, stream=None, ef_errors=None, scale=0.0):
    """ring_quantizer.py:25-49: the lossy running SUM; ef_errors[u][i] (updated in place)
    enables error feedback (:33-38)."""
    U = len(user_grads)
    prev = [None] * len(codecs)
    for u in range(U):
        for i, c in enumerate(codecs):
            g = np.asarray(user_grads[u][i], dtype=np.float32)
            if u != 0:
                g = g + prev[i]
            if ef_errors is not None:
                g = g + np.float32(scale) * ef_errors[u][i]
            d = c.decompress(c.compress(g, stream))
            if ef_errors is not None:
                ef_errors[u][i] = g - d
            prev[i] = d
    return prev
