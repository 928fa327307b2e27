#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED Python reference.

Run in the authoring container only (needs /root/reference, which does not exist
on the GPU box):

    python tests/golden/make_golden.py

The reference is imported from /root/reference with cwd set there (its codebook
path is cwd-relative, compressors/nearest_neighbor_compressor.py:50).  Each case
records the seeded inputs' recipe, the uniform draws the reference consumed
(torch.manual_seed(seed) + torch.rand in reference call order) and the
reference's outputs.  The script also checks oracle/gq_oracle.py against every
case while generating (this is what "pins" the oracle) and exits non-zero on any
mismatch.

Inputs are regenerated in the tests from numpy's frozen legacy RandomState
stream, so fixtures only hold draws + outputs.
"""
import hashlib
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
os.chdir(REF)

from compressors import (IdenticalCompressor, NearestNeighborCompressor,  # noqa: E402
                         ProbabilisticVectorCompressor, QSGDCompressor, ResidualCompressor,
                         SignSGDCompressor, TopKSparsificationCompressor)
from quantizers import Quantizer  # noqa: E402

from oracle import gq_oracle as O  # noqa: E402

torch.set_num_threads(1)
FAIL = []


def args(**kw):
    base = dict(c_dim=16, k_bit=8, n_bit=6, no_cuda=True, random=True, cr=256, ef=False,
                two_phase=False, mode="ps", scale="exp", num_users=8)
    base.update(kw)
    return SimpleNamespace(**base)


def gen_input(seed, n, kind="normal"):
    """Frozen-stream synthetic gradient (also used verbatim by tests/util.py)."""
    rs = np.random.RandomState(seed)
    if kind == "normal":
        return (rs.standard_normal(n) * 0.01).astype(np.float32)
    if kind == "heavy":      # heavy-tailed magnitudes over many octaves
        return (rs.standard_normal(n) * np.exp(rs.standard_normal(n) * 3.0) * 1e-3).astype(np.float32)
    if kind == "repeat16":   # every 16-chunk identical -> all u equal -> lb == ub branch
        return np.tile((rs.standard_normal(16) * 0.01).astype(np.float32), n // 16)
    if kind == "zeros_mixed":  # some all-zero chunks (QSGD 0/0 edge)
        x = (rs.standard_normal(n) * 0.01).astype(np.float32)
        x[: n // 4] = 0.0
        return x
    raise ValueError(kind)


def check(name, ok):
    if not ok:
        FAIL.append(name)
        print("  MISMATCH:", name)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def codebook(d, K):
    return O.normalize(O.fvecs_read(
        os.path.join(REF, "codebooks/learned_codebook/angular_dim_%d_Ks_%d.fvecs" % (d, K))))[1]


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print("wrote %-34s %8.1f KB" % (name + ".npz", os.path.getsize(path) / 1024))


# ----------------------------------------------------------------------------
def hsq_case(name, shape, c_dim, k_bit, n_bit, random, seed, kind="normal"):
    size = int(np.prod(shape))
    a = args(c_dim=c_dim, k_bit=k_bit, n_bit=n_bit, random=random)
    comp = NearestNeighborCompressor(size, torch.Size(shape), a)
    x = gen_input(seed, size, kind)
    torch.manual_seed(seed)
    sig = comp.compress(torch.from_numpy(x).view(shape))
    dec = comp.decompress(sig).numpy()
    nchunks = size // comp.dim
    torch.manual_seed(seed)
    draws = torch.rand(nchunks).numpy()
    out = dict(shape=np.array(shape), c_dim=c_dim, k_bit=k_bit, n_bit=n_bit, random=int(random),
               seed=seed, kind=kind, dim=comp.dim, codes=sig[1].numpy().astype(np.int32),
               decoded=dec.reshape(-1), draws=draws, input_sha=sha(x))
    if n_bit != 32:
        lb, ub, l = sig[0]
        out.update(lb=np.float32(lb.item()), ub=np.float32(ub.item()), l=l.numpy().astype(np.int32))
    else:
        out.update(u=sig[0].numpy())
    # pin the oracle
    cb = codebook(comp.dim, comp.K)
    check(name + ":codebook", np.array_equal(cb, comp.codewords.numpy()))
    oc = O.HSQ(size, shape, cb, n_bit, random)
    osig = oc.compress(x, O.UniformStream(draws))
    check(name + ":codes", np.array_equal(osig[1], out["codes"]))
    if n_bit != 32:
        check(name + ":lb/ub", osig[0][0] == out["lb"] and osig[0][1] == out["ub"])
        check(name + ":l", np.array_equal(osig[0][2], out["l"]))
    else:
        check(name + ":u", np.array_equal(osig[0], out["u"]))
    check(name + ":decoded", np.array_equal(oc.decompress(osig).reshape(-1), out["decoded"]))
    save(name, **out)


def qsgd_case(name, shape, c_dim, n_bit, random, seed, kind="normal"):
    size = int(np.prod(shape))
    a = args(c_dim=c_dim, n_bit=n_bit, random=random)
    comp = QSGDCompressor(size, torch.Size(shape), a)
    x = gen_input(seed, size, kind)
    torch.manual_seed(seed)
    norm, signs, l = comp.compress(torch.from_numpy(x).view(shape))
    dec = comp.decompress([norm, signs, l]).numpy()
    torch.manual_seed(seed)
    draws = torch.rand(size).numpy()
    out = dict(shape=np.array(shape), c_dim=c_dim, n_bit=n_bit, random=int(random), seed=seed,
               kind=kind, dim=comp.dim, norm=norm.numpy().reshape(-1),
               signs=signs.numpy().reshape(-1).astype(np.uint8),
               l=l.numpy().reshape(-1).astype(np.int32), decoded=dec.reshape(-1),
               draws_sha=sha(draws), input_sha=sha(x))
    oc = O.QSGD(size, shape, c_dim, n_bit, random)
    check(name + ":dim", oc.dim == comp.dim)
    osig = oc.compress(x, O.UniformStream(draws))
    check(name + ":norm", np.array_equal(osig[0], out["norm"]))
    check(name + ":signs", np.array_equal(osig[1], out["signs"]))
    check(name + ":l", np.array_equal(osig[2], out["l"]))
    od = oc.decompress(osig).reshape(-1)
    check(name + ":decoded", np.array_equal(od, out["decoded"], equal_nan=True))
    save(name, **out)


def sign_topk_case(seed=5):
    shape = (64, 3, 3, 16)
    size = int(np.prod(shape))
    x = gen_input(seed, size, "normal")
    x[::7] = 0.0  # exact zeros: sign(0) = 0
    s = SignSGDCompressor(size, torch.Size(shape), args()).compress(torch.from_numpy(x).view(shape)).numpy()
    check("sign", np.array_equal(O.sign(x), s.reshape(-1)))
    out = dict(shape=np.array(shape), seed=seed, sign=s.reshape(-1).astype(np.int8), input_sha=sha(x))
    for cr in (100, 256, 3):
        comp = TopKSparsificationCompressor(size, torch.Size(shape), args(cr=cr))
        t = comp.decompress(comp.compress(torch.from_numpy(x).view(shape))).numpy().reshape(-1)
        o = O.topk(x, size // cr)
        check("topk cr=%d values" % cr, np.array_equal(o, t))
        check("topk cr=%d signed zeros" % cr, np.array_equal(np.signbit(o), np.signbit(t)))
        nz = np.flatnonzero(t)
        out["topk_idx_cr%d" % cr] = nz.astype(np.int32)
    save("sign_topk", **out)


class _P:
    """Stand-in for torch.nn.Parameter: the quantizers only touch .grad/.shape/.flatten()."""


class _runnable_pvc:
    """What it takes to EXECUTE the reference's ProbabilisticVectorCompressor, without
    touching its source:
      * it opens './codebook/angular_dim_{d}_Ks_{K}.fvecs' (probabilistic_vector_compressor.py:26), a
        directory the checkout does not have -> run with cwd = a scratch directory whose `codebook`
        entry links to the reference's codebooks/learned_codebook (the files HSQ loads);
      * `torch.argmin(comp, dim=1) + 1` on a bool tensor (:57-58) raises on every PyTorch >= 1.2.  On the
        ByteTensor `comp` of the PyTorch it was written for, argmin returned the LAST index of the
        minimum, so argmin + 1 = the first index whose cumulative probability reaches r - 1e-5 -- and
        = K (out of range, the reference's gather then fails) for rows whose first probability already
        reaches it.  Cast-to-uint8 with today's first-index argmin gives code == 1 for every row, the
        last-index emulation crashes on ~1/K of the rows (tests/golden/pvc_pin_attempt.txt).  Here
        torch.argmin is replaced, for the duration of compress(), by `first True index - 1`: every
        other line (p = pinv(C^T) v, L1 norms, probabilities, cumsum, threshold, u, norm quantizer,
        decode, the residual loop) runs as shipped."""

    def __init__(self):
        import tempfile
        self.dir = tempfile.mkdtemp(prefix="gq_pvc_")
        os.symlink(os.path.join(REF, "codebooks", "learned_codebook"), os.path.join(self.dir, "codebook"))
        os.symlink(os.path.join(REF, "codebooks"), os.path.join(self.dir, "codebooks"))

    def __enter__(self):
        self.cwd = os.getcwd()
        os.chdir(self.dir)
        self.argmin = torch.argmin
        torch.argmin = lambda t, dim=None: torch.argmax(t.to(torch.uint8), dim=dim) - 1
        return self

    def __exit__(self, *exc):
        torch.argmin = self.argmin
        os.chdir(self.cwd)


def _sig_arrays(prefix, sig, n_bit, out):
    if n_bit != 32:
        lb, ub, l = sig[0]
        out[prefix + "lb"] = np.float32(lb.item())
        out[prefix + "ub"] = np.float32(ub.item())
        out[prefix + "l"] = l.numpy().astype(np.int32)
    else:
        out[prefix + "u"] = sig[0].numpy()
    out[prefix + "codes"] = sig[1].numpy().astype(np.int32)


def pvc_case(name, n_chunks, d, k_bit, n_bit, seed, residual=False):
    """ProbabilisticVectorCompressor / ResidualCompressor through the reference's own code."""
    size = n_chunks * d
    shape = (n_chunks, d)
    a = args(c_dim=d, k_bit=k_bit, n_bit=n_bit)
    x = gen_input(seed, size)
    with _runnable_pvc():
        np.random.seed(seed)           # scipy.stats.ortho_group (K == d) draws from numpy's global state
        comp = (ResidualCompressor if residual else ProbabilisticVectorCompressor)(size, torch.Size(shape), a)
        torch.manual_seed(seed)
        sig = comp.compress(torch.from_numpy(x.copy()).view(shape))
        dec = comp.decompress(sig).numpy().reshape(-1)
        after = torch.rand(1).item()
    pvc = comp.compressors[1] if residual else comp
    cb = pvc.codewords.numpy()
    n_draws = n_chunks * ((2 if n_bit != 32 else 1) + ((1 if n_bit != 32 else 0) if residual else 0))
    torch.manual_seed(seed)
    draws = torch.rand(n_draws).numpy()
    check(name + ":draw count", torch.rand(1).item() == after)
    out = dict(n_chunks=n_chunks, d=d, k_bit=k_bit, n_bit=n_bit, seed=seed, residual=int(residual),
               codewords=cb, dagger=pvc.c_dagger.numpy(), decoded=dec, draws=draws, input_sha=sha(x))
    stream = O.UniformStream(draws)
    if residual:
        _sig_arrays("s1_", sig[0], n_bit, out)
        _sig_arrays("s2_", sig[1], n_bit, out)
        check(name + ":stage-1 codebook", np.array_equal(comp.compressors[0].codewords.numpy(), cb))
        oc = O.Residual(size, shape, cb, n_bit, True)
        oc.stages[1].dagger = out["dagger"]
        osig = oc.compress(x, stream)
        pairs = (("s1_", osig[0]), ("s2_", osig[1]))
    else:
        _sig_arrays("", sig, n_bit, out)
        oc = O.PVC(size, shape, cb, n_bit, True)
        oc.dagger = out["dagger"]
        osig = oc.compress(x, stream)
        pairs = (("", osig),)
    check(name + ":pinv", np.array_equal(np.linalg.pinv(cb.T).astype(np.float32), out["dagger"]))
    for pre, s in pairs:
        check(name + ":" + pre + "codes", np.array_equal(s[1], out[pre + "codes"]))
        if n_bit != 32:
            check(name + ":" + pre + "lb/ub", s[0][0] == out[pre + "lb"] and s[0][1] == out[pre + "ub"])
            check(name + ":" + pre + "l", np.array_equal(s[0][2], out[pre + "l"]))
        else:
            check(name + ":" + pre + "u", np.array_equal(s[0], out[pre + "u"]))
    check(name + ":decoded", np.array_equal(oc.decompress(osig).reshape(-1), dec))
    check(name + ":codes are not degenerate", len(np.unique(out["s2_codes" if residual else "codes"])) > 4)
    save(name, **out)


def pvc_pin_attempt():
    """Record what the un-patched alternatives do (why PVC needs the one-line reading above)."""
    lines = []
    a = args(c_dim=16, k_bit=4, n_bit=32)
    x = gen_input(3, 16 * 4000)
    with _runnable_pvc() as ctx:
        np.random.seed(5)
        c = ProbabilisticVectorCompressor(16 * 4000, torch.Size((4000, 16)), a)
        first_true = torch.argmin
        torch.argmin = ctx.argmin
        try:
            torch.manual_seed(7)
            c.compress(torch.from_numpy(x))
            lines.append("as shipped: ran")
        except Exception as e:  # noqa: BLE001
            lines.append("as shipped: %s: %s" % (type(e).__name__, str(e).splitlines()[0]))
        torch.argmin = lambda t, dim=None: ctx.argmin(t.to(torch.uint8), dim=dim)
        torch.manual_seed(7)
        _, codes = c.compress(torch.from_numpy(x))
        lines.append("bool cast to uint8, today's first-index argmin: codes take the values %s on 4000 chunks"
                     % sorted(set(codes.numpy().tolist())))

        def argmin_last(t, dim=None):
            t8 = t.to(torch.uint8)
            return (t8.shape[1] - 1) - ctx.argmin(torch.flip(t8, dims=[1]), dim=1)
        torch.argmin = argmin_last
        try:
            torch.manual_seed(7)
            c.compress(torch.from_numpy(x))
            lines.append("last-index argmin (old ByteTensor behaviour): ran")
        except Exception as e:  # noqa: BLE001
            lines.append("last-index argmin (old ByteTensor behaviour): %s: %s" % (type(e).__name__, str(e).splitlines()[0]))
        torch.argmin = first_true
        vec = torch.from_numpy(x).view(-1, 16)
        p = torch.mm(c.c_dagger, vec.transpose(0, 1)).transpose(0, 1)
        prob = torch.abs(p) / torch.norm(p, p=1, dim=1, keepdim=True)
        torch.manual_seed(7)
        r = torch.rand(4000)
        comp = torch.cumsum(prob, dim=1) >= r.view(-1, 1).expand_as(prob) - 1e-5
        lines.append("rows whose first cumulative probability already reaches r - 1e-5 (argmin + 1 == K there): %d of 4000"
                     % int(comp.all(dim=1).sum()))
        oc, ou = O.pvc_search(x.reshape(-1, 16), c.c_dagger.numpy(), r.numpy())
        lines.append("oracle codes vs first True index of the reference's own `comp`: %d mismatches"
                     % int((torch.from_numpy(oc) != torch.argmax(comp.to(torch.uint8), dim=1)).sum()))
    with open(os.path.join(HERE, "pvc_pin_attempt.txt"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    print("\n".join("  pvc: " + ln for ln in lines))


def quantizer_case(name, mode, shapes, U, seed, ef=False, quant="hsq", epoch=1, iters=1, **kw):
    a = args(mode=mode, num_users=U, ef=ef, **kw)
    Comp = {"hsq": NearestNeighborCompressor, "qsgd": QSGDCompressor, "sign": SignSGDCompressor,
            "topk": TopKSparsificationCompressor}[quant]
    params = [torch.nn.Parameter(torch.zeros(s)) for s in shapes]
    q = Quantizer(Comp, params, a)
    sizes = [int(np.prod(s)) for s in shapes]
    torch.manual_seed(seed)
    results = []
    for it in range(iters):
        grads = [[gen_input(seed * 1000 + it * 100 + u * 10 + i, n) for i, n in enumerate(sizes)]
                 for u in range(U)]
        for u in range(U):
            for p, g, s in zip(params, grads[u], shapes):
                p.grad = torch.from_numpy(g.copy()).view(s)
            q.record(u, epoch=epoch)
        q.apply()
        results.append([p.grad.data.numpy().reshape(-1).copy() for p in params])
    # the stream of uniforms consumed, replayed
    gen_state_after = torch.rand(1).item()
    torch.manual_seed(seed)
    total = 0
    # upper bound on draws: every compressed tensor, every user, every iter
    per_pass = 0
    for c in q.compressors:
        if isinstance(c, NearestNeighborCompressor):
            per_pass += c.size // c.dim
        elif isinstance(c, QSGDCompressor):
            per_pass += c.size
    total = per_pass * (U + (1 if a.two_phase else 0)) * iters      # --two-phase: one more compression in apply()
    draws = torch.rand(total).numpy() if total else np.zeros(0, np.float32)
    check(name + ":draw count", total == 0 or torch.rand(1).item() == gen_state_after)

    # oracle replay
    codecs = []
    for c, s, n in zip(q.compressors, shapes, sizes):
        if isinstance(c, NearestNeighborCompressor):
            codecs.append(O.HSQ(n, s, codebook(c.dim, c.K), a.n_bit, a.random))
        elif isinstance(c, QSGDCompressor):
            codecs.append(O.QSGD(n, s, a.c_dim, a.n_bit, a.random))
        elif isinstance(c, SignSGDCompressor):
            codecs.append(O.Sign(n, s))
        elif isinstance(c, TopKSparsificationCompressor):
            codecs.append(O.TopK(n, s, a.cr))
        else:
            codecs.append(O.Identity())
    stream = O.UniformStream(draws)
    errs = [[np.zeros(s, np.float32) for s in shapes] for _ in range(U)] if ef else None
    serrs = [np.zeros(s, np.float32) for s in shapes] if (ef and a.two_phase) else None
    for it in range(iters):
        grads = [[gen_input(seed * 1000 + it * 100 + u * 10 + i, n).reshape(s)
                  for i, (n, s) in enumerate(zip(sizes, shapes))] for u in range(U)]
        if mode == "ps":
            og = O.ps_step(codecs, grads, stream, errs, O.ps_scale(epoch, a.scale), a.two_phase, serrs)
        else:
            og = O.ring_step(codecs, grads, stream, errs, O.ps_scale(epoch, a.scale))
        for i in range(len(shapes)):
            ref = results[it][i]
            got = og[i].reshape(-1)
            exact = np.array_equal(ref, got)
            rel = np.abs(ref - got).max() / max(np.abs(ref).max(), 1e-30)
            check("%s:it%d:tensor%d (exact=%s rel=%.2e)" % (name, it, i, exact, rel), rel <= 1e-6)
    out = dict(mode=mode, U=U, seed=seed, ef=int(ef), two_phase=int(a.two_phase), quant=quant, epoch=epoch, iters=iters,
               n_tensors=len(shapes), draws_sha=sha(draws), n_draws=total,
               c_dim=a.c_dim, k_bit=a.k_bit, n_bit=a.n_bit, cr=a.cr)
    for i, s in enumerate(shapes):
        out["shape%d" % i] = np.array(s)
        for it in range(iters):
            out["grad_it%d_t%d" % (it, i)] = results[it][i]
    save(name, **out)


def mm_claim():
    """CPU torch.mm == ascending-j fmaf chain, bit for bit."""
    for d, K in ((8, 256), (16, 256), (32, 256), (16, 4096)):
        cb = codebook(d, K)
        x = gen_input(99 + d, 20000 * d).reshape(-1, d)
        p = torch.mm(torch.from_numpy(cb), torch.from_numpy(x).transpose(0, 1)).transpose(0, 1).numpy()
        o = O.hsq_scores(x, cb)
        neq = int((p != o).sum())
        print("  torch.mm vs fmaf chain d=%d K=%d: %d / %d elements differ" % (d, K, neq, p.size))
        codes_t = np.abs(p).argmax(1)
        codes_o, _ = O.hsq_search(x, cb)
        check("mm claim codes d=%d K=%d" % (d, K), np.array_equal(codes_t, codes_o))


if __name__ == "__main__":
    print("reference: %s   torch %s   numpy %s" % (REF, torch.__version__, np.__version__))
    only = sys.argv[1:]          # optional: regenerate only the cases whose name contains one of these
    if only:
        _save = save

        def save(name, **arrs):  # noqa: F811
            if any(o in name for o in only):
                _save(name, **arrs)
    mm_claim()
    hsq_case("hsq_d16_k256_n6", (64, 256), 16, 8, 6, True, 11)
    hsq_case("hsq_d16_k256_n6_fcn2", (10, 256), 16, 8, 6, True, 12)
    hsq_case("hsq_d16_k256_n6_heavy", (32, 16, 3, 3), 16, 8, 6, True, 13, "heavy")
    hsq_case("hsq_d16_k256_n6_norandom", (64, 64), 16, 8, 6, False, 14)
    hsq_case("hsq_d16_k256_n32", (64, 64), 16, 8, 32, True, 15)
    hsq_case("hsq_d16_k256_n8", (64, 64), 16, 8, 8, True, 16)
    hsq_case("hsq_d16_k256_lbub_equal", (64, 64), 16, 8, 6, True, 17, "repeat16")
    hsq_case("hsq_d8_k256_n6", (128, 64), 8, 8, 6, True, 18)
    hsq_case("hsq_d32_k256_n6", (128, 64), 32, 8, 6, True, 19)
    hsq_case("hsq_d16_k4096_n6", (128, 64), 16, 12, 6, True, 20)
    hsq_case("hsq_conv1_d16_k256_n6", (64, 3, 3, 3), 16, 8, 6, True, 21)
    hsq_case("hsq_d16to24_k256_n6", (45, 24), 16, 8, 6, True, 22)  # 1080 -> dim 24
    qsgd_case("qsgd_d128_n2", (64, 256), 128, 2, True, 31)
    qsgd_case("qsgd_d128to192_n2", (64, 3, 3, 3), 128, 2, True, 32)
    qsgd_case("terngrad_n1", (64, 64), 0, 1, True, 33)
    qsgd_case("qsgd_d128_n2_zero_chunks", (64, 128), 128, 2, True, 34, "zeros_mixed")
    qsgd_case("qsgd_d128_n4_norandom", (64, 128), 128, 4, False, 35)
    sign_topk_case()
    fcn = [(256, 784), (256,), (10, 256), (10,)]
    quantizer_case("ps_fcn_hsq_u8", "ps", fcn, 8, 41)
    quantizer_case("ring_fcn_hsq_u4", "ring", fcn, 4, 42)
    quantizer_case("ps_fcn_hsq_u4_ef", "ps", fcn, 4, 43, ef=True, iters=2)
    small = [(64, 128), (100,), (32, 64)]
    quantizer_case("ps_small_qsgd_u4", "ps", small, 4, 44, quant="qsgd", c_dim=128, n_bit=2)
    quantizer_case("ring_small_qsgd_u3", "ring", small, 3, 45, quant="qsgd", c_dim=128, n_bit=2)
    quantizer_case("ps_small_sign_u4", "ps", small, 4, 46, quant="sign")
    quantizer_case("ps_small_topk_u4", "ps", small, 4, 47, quant="topk", cr=100)
    quantizer_case("ps_small_hsq_u4_2p", "ps", small, 4, 48, two_phase=True, iters=2)
    quantizer_case("ps_small_hsq_u3_2p_ef", "ps", small, 3, 49, ef=True, two_phase=True, iters=2)
    quantizer_case("ring_small_hsq_u3_ef", "ring", small, 3, 50, ef=True, iters=2)
    pvc_pin_attempt()
    pvc_case("pvc_kd16_n32", 2000, 16, 4, 32, 61)           # K == d: random orthogonal basis (ortho_group)
    pvc_case("pvc_d16_k256_n6", 2000, 16, 8, 6, 62)         # learned codebook, quantized norms
    pvc_case("pvc_d8_k256_n6", 3000, 8, 8, 6, 63)
    pvc_case("residual_d16_k256_n6", 2000, 16, 8, 6, 64, residual=True)
    if FAIL:
        print("\nORACLE != REFERENCE in %d checks:" % len(FAIL))
        for f in FAIL:
            print("  ", f)
        sys.exit(1)
    print("\noracle pinned: every case bit-identical to the live reference")
