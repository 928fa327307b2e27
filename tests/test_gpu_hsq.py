"""GPU parity tests for the HSQ path (through the C ABI, via the reference-named classes).
Bit-exact bar: codes, u, lb/ub, l (given the same uniforms) and -- because decode is
one fp32 multiply -- the decoded gradient."""
import numpy as np
import pytest
import torch

import gq_b200
from gq_b200 import _lib
from oracle import gq_oracle as O
from util import codebook, gen_input, golden, golden_names, make_args, torch_uniform_stream

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _t(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


@pytest.mark.parametrize("algo", [_lib.ALGO_EXACT, _lib.ALGO_AUTO])
@pytest.mark.parametrize("name", golden_names("hsq_"))
def test_hsq_golden(name, algo):
    g = golden(name)
    shape = tuple(int(x) for x in g["shape"])
    size = int(np.prod(shape))
    a = make_args(c_dim=int(g["c_dim"]), k_bit=int(g["k_bit"]), n_bit=int(g["n_bit"]),
                  random=bool(g["random"]), hsq_algo=algo)
    c = gq_b200.NearestNeighborCompressor(size, torch.Size(shape), a)
    assert c.dim == int(g["dim"])
    x = _t(gen_input(int(g["seed"]), size, str(g["kind"]))).view(shape)
    sig = c.compress(x, uniforms=g["draws"])
    assert sig[1].dtype == (torch.uint8 if int(g["k_bit"]) <= 8 else torch.int32)
    assert np.array_equal(sig[1].cpu().numpy().astype(np.int32), g["codes"])
    if int(g["n_bit"]) != 32:
        lb, ub, l = sig[0]
        assert l.dtype == torch.int32
        assert np.float32(lb.item()) == g["lb"] and np.float32(ub.item()) == g["ub"]
        assert np.array_equal(l.cpu().numpy(), g["l"])
    else:
        assert np.array_equal(sig[0].cpu().numpy(), g["u"])
    dec = c.decompress(sig)
    assert tuple(dec.shape) == shape
    assert np.array_equal(dec.cpu().numpy().reshape(-1), g["decoded"])


@pytest.mark.parametrize("d,K", [(16, 256), (8, 256), (32, 256), (16, 4096), (24, 256), (12, 256), (64, 256)])
@pytest.mark.parametrize("kind", ["normal", "heavy"])
def test_hsq_search_vs_oracle(d, K, kind):
    n_chunks = 40000 + 37  # ragged against every tile size
    x = gen_input(1000 + d + K, n_chunks * d, kind)
    cb = codebook(d, K)
    oc, ou = O.hsq_search(x.reshape(-1, d), cb)
    for algo in (_lib.ALGO_EXACT, _lib.ALGO_AUTO):
        a = make_args(c_dim=d, k_bit=int(np.log2(K)), n_bit=32, hsq_algo=algo)
        c = gq_b200.NearestNeighborCompressor(n_chunks * d, torch.Size((n_chunks, d)), a)
        u, codes = c.compress(_t(x))
        assert np.array_equal(codes.cpu().numpy().astype(np.int32), oc), (d, K, algo)
        assert np.array_equal(u.cpu().numpy(), ou), (d, K, algo)


@pytest.mark.parametrize("d", [16, 8, 32])
def test_hsq_adversarial_ties_and_specials(d):
    """Exact ties (first index wins), zero chunks, codeword inputs, huge/tiny magnitudes -- at every
    chunk dimension the tcgen05 encode handles."""
    K = 256
    cb = codebook(d, K)
    rows = []
    rows.append(np.zeros(d, np.float32))                              # all scores 0 -> code 0
    rows.append(cb[7] * np.float32(3.0))                              # a codeword itself
    rows.append(-cb[200] * np.float32(1e-20))                         # tiny, negative projection
    rows.append(cb[5] * np.float32(1e20))                             # huge
    rows.append((cb[9] + cb[10]).astype(np.float32))                  # near tie between 9 and 10
    rows.append((cb[10] + cb[9]).astype(np.float32))
    rs = np.random.RandomState(5)
    for _ in range(200):                                              # sparse one-hot style chunks
        r = np.zeros(d, np.float32)
        r[rs.randint(d)] = rs.standard_normal()
        rows.append(r)
    for _ in range(300):                                              # low-precision values: many exact ties
        rows.append((rs.randint(-2, 3, size=d)).astype(np.float32))
    for sc in (1e-18, 1e-30, 1e-36, 3e-39, 1e-42):                    # down into the denormals, where the
        for _ in range(40):                                           # tensor core may flush operands
            rows.append((rs.standard_normal(d) * sc).astype(np.float32))
    for _ in range(40):                                               # mixed: one normal element, rest denormal
        r = (rs.standard_normal(d) * 1e-40).astype(np.float32)
        r[rs.randint(d)] = np.float32(2e-38)
        rows.append(r)
    rows.append(np.full(d, -0.0, np.float32))                         # all negative zeros
    x = np.stack(rows * 8).astype(np.float32)
    oc, ou = O.hsq_search(x, cb)
    for algo in (_lib.ALGO_EXACT, _lib.ALGO_AUTO):
        a = make_args(c_dim=d, n_bit=32, hsq_algo=algo)
        c = gq_b200.NearestNeighborCompressor(x.size, torch.Size(x.shape), a)
        u, codes = c.compress(_t(x))
        assert np.array_equal(codes.cpu().numpy().astype(np.int32), oc), algo
        assert np.array_equal(u.cpu().numpy(), ou), algo


def _raw_search(x, cb, algo):
    n_chunks, d = x.shape
    xt, cbt = _t(x), _t(cb)
    codes = torch.full((n_chunks,), 255, dtype=torch.uint8, device=DEV)
    u = torch.zeros(n_chunks, device=DEV)
    seg = torch.tensor([0, n_chunks], dtype=torch.int64, device=DEV)
    ws = torch.empty(1 << 20, dtype=torch.uint8, device=DEV)
    _lib.call("gq_hsq_search", xt.data_ptr(), n_chunks, d, cbt.data_ptr(), cb.shape[0], codes.data_ptr(), 1,
              u.data_ptr(), seg.data_ptr(), 1, None, ws.data_ptr(), ws.numel(), algo, _lib.stream())
    torch.cuda.synchronize()
    return codes.cpu().numpy().astype(np.int32), u.cpu().numpy()


@pytest.mark.parametrize("scale", ["x3.7", "rows", "tiny"])
def test_hsq_search_exact_for_codebooks_that_are_not_unit_norm(scale):
    """The tensor-core filter's margin follows the largest codeword norm, so the C entry point stays
    bit-exact when the caller's codebook is not normalised (the reference always normalises)."""
    rs = np.random.RandomState(11)
    cb = codebook(16, 256).copy()
    if scale == "x3.7":
        cb *= np.float32(3.7)
    elif scale == "rows":
        cb *= rs.uniform(0.05, 20.0, size=(256, 1)).astype(np.float32)
    else:
        cb *= np.float32(1e-12)
    x = gen_input(77, 30011 * 16, "normal").reshape(-1, 16)
    oc, ou = O.hsq_search(x, cb)
    for algo in (_lib.ALGO_EXACT, _lib.ALGO_AUTO):
        gc, gu = _raw_search(x, cb, algo)
        assert np.array_equal(gc, oc), (scale, algo)
        assert np.array_equal(gu, ou), (scale, algo)


def test_hsq_nonfinite_inputs_do_not_hang_and_match_nan_rule():
    d, K = 16, 256
    x = gen_input(77, 256 * d).reshape(-1, d).copy()
    x[3, 2] = np.inf
    x[10, 0] = np.nan
    x[20, :] = -np.inf
    cb = codebook(d, K)
    oc, ou = O.hsq_search(x, cb)
    for algo in (_lib.ALGO_EXACT, _lib.ALGO_AUTO):
        a = make_args(n_bit=32, hsq_algo=algo)
        c = gq_b200.NearestNeighborCompressor(x.size, torch.Size(x.shape), a)
        u, codes = c.compress(_t(x))
        got = codes.cpu().numpy().astype(np.int32)
        finite = np.isfinite(x).all(1)
        assert np.array_equal(got[finite], oc[finite])
        assert np.array_equal(got, oc), "non-finite rows follow torch.argmax's NaN-is-max rule"


def test_norm_quantizer_class_and_levels():
    a = make_args(n_bit=6)
    psc = gq_b200.ProbabilisticScalarCompressor(6, a)
    u = gen_input(3, 4096)
    r = torch_uniform_stream(3, 4096)
    lb, ub, l = psc.compress(_t(u), uniforms=r)
    olb, oub, ol, _ = O.psc_compress(u, 6, True, r)
    assert np.float32(lb.item()) == olb and np.float32(ub.item()) == oub
    assert np.array_equal(l.cpu().numpy(), ol)
    assert int(l.min()) == 0 and int(l.max()) == 64          # 2^n + 1 levels
    back = psc.decompress((lb, ub, l))
    assert np.array_equal(back.cpu().numpy(), O.psc_decompress(olb, oub, ol, 6))
    # lb == ub -> zeros, handled on the device (no sync)
    lb, ub, l = psc.compress(torch.full((100,), 0.25, device=DEV), uniforms=r)
    assert not l.any() and lb.item() == ub.item() == 0.25


def test_philox_stream_is_uniform_and_reproducible():
    a = make_args(n_bit=6)
    psc = gq_b200.ProbabilisticScalarCompressor(6, a)
    # u uniform in [0,1): scaled*64 has a uniform fractional part; E[l] == u*64 (unbiased)
    u = torch.rand(1 << 20, device=DEV)
    u[0], u[1] = 0.0, 1.0
    torch.manual_seed(123)
    _, _, l1 = psc.compress(u)
    torch.manual_seed(123)
    _, _, l2 = psc.compress(u)
    assert torch.equal(l1, l2)
    _, _, l3 = psc.compress(u)
    assert not torch.equal(l1, l3)                            # offset advanced
    err = (l1.float() / 64.0 - u).double()
    assert abs(err.mean().item()) < 2e-5                      # unbiased (sigma/sqrt(n) ~ 7e-6)
    assert err.abs().max().item() <= 1.0 / 64 + 1e-6


def test_hsq_segment_table_matches_per_tensor_calls():
    """One fused call over 5 tensors == 5 per-tensor calls (lb/ub are per tensor)."""
    from gq_b200.quantizers.fused import FusedPlan
    shapes = [(64, 64), (7,), (32, 16, 3, 3), (1024,), (100, 48), (16, 80)]
    a = make_args(num_users=1)
    plan = FusedPlan(gq_b200.NearestNeighborCompressor, shapes, a, torch.device(DEV), 1)
    xs = [gen_input(50 + i, int(np.prod(s))).reshape(s) for i, s in enumerate(shapes)]
    plan.gather([_t(x) for x in xs])
    n_draws = sum(x.size // 16 for x in xs if x.size > 1000)
    stream = torch_uniform_stream(9, n_draws)
    parts, used = plan.split_uniform_stream(stream)
    assert used == n_draws
    plan.encode(0, uniforms=parts)
    out = plan.decode(mean=True)
    torch.cuda.synchronize()
    s = O.UniformStream(stream)
    for i, (x, shp) in enumerate(zip(xs, shapes)):
        got = plan.view(i, out).cpu().numpy()
        if x.size <= 1000:
            assert np.array_equal(got, x)
            continue
        oc = O.HSQ(x.size, shp, codebook(16, 256), 6, True)
        ref = oc.decompress(oc.compress(x, s))
        assert np.array_equal(got, ref), i


def test_host_roundtrip_entry():
    d, K = 16, 256
    n_chunks = 5000
    x = gen_input(8, n_chunks * d)
    cb = _t(codebook(d, K))
    seg = torch.tensor([0, 1200, n_chunks], dtype=torch.int64, device=DEV)
    need = _lib.value("gq_hsq_host_scratch_bytes", n_chunks, d, K, 2)
    scratch = torch.empty(need, dtype=torch.uint8, device=DEV)
    hx = torch.from_numpy(x).pin_memory()
    hy = torch.empty_like(hx).pin_memory()
    _lib.call("gq_hsq_roundtrip_host", hx.data_ptr(), hy.data_ptr(), n_chunks, d, cb.data_ptr(), K,
              seg.data_ptr(), 2, 6, 0, 0, 0, scratch.data_ptr(), need, _lib.ALGO_AUTO, _lib.stream())
    ref = []
    for lo, hi in ((0, 1200), (1200, n_chunks)):
        part = x[lo * d:hi * d]
        oc = O.HSQ(part.size, (part.size,), codebook(d, K), 6, False)
        ref.append(oc.decompress(oc.compress(part)))
    assert np.array_equal(hy.numpy(), np.concatenate(ref))


def test_errors_are_loud():
    cb = _t(codebook(16, 256))
    x = torch.zeros(160, device=DEV)
    with pytest.raises(_lib.GQError):
        _lib.call("gq_hsq_encode", x.data_ptr(), 10, 16, cb.data_ptr(), 256, None, 0, 6, 1, None, 0, 0,
                  None, 1, None, 1, None, None, None, 0, 0, _lib.stream())
    with pytest.raises(_lib.GQError):
        gq_b200.NearestNeighborCompressor(4096, (64, 64), make_args()).compress(torch.zeros(4096))


def test_codebook_generator_on_gpu_and_k65536_search():
    """The generator's HSQ objective runs this package's search kernel; a synthesised K = 2^16
    codebook (BASELINE config 4: no reference file exists) goes through the exact search and
    equals the oracle on the same array."""
    from gq_b200 import codebook_generator as G
    cb = G.train_codebook(16, 512, train_size=20000, iter=4, objective="hsq")
    assert cb.shape == (512, 16) and np.isfinite(cb).all()
    assert abs(np.linalg.norm(cb, axis=1) - 1).max() < 1e-5
    rnd = np.random.RandomState(0).standard_normal((512, 16)).astype(np.float32)
    assert G.quantisation_error(cb, n=20000) < G.quantisation_error(rnd, n=20000)
    big = O.normalize(np.random.RandomState(4).standard_normal((65536, 16)).astype(np.float32))[1]
    x = gen_input(77, 16 * 3000)
    n = 3000
    codes = torch.empty(n, dtype=torch.int32, device=DEV)
    u = torch.empty(n, device=DEV)
    seg = torch.tensor([0, n], dtype=torch.int64, device=DEV)
    ws = torch.empty(1 << 16, dtype=torch.uint8, device=DEV)
    cbt = torch.from_numpy(big).to(DEV)
    xt = torch.from_numpy(x).to(DEV)
    _lib.call("gq_hsq_search", xt.data_ptr(), n, 16, cbt.data_ptr(), 65536, codes.data_ptr(), 4, u.data_ptr(),
              seg.data_ptr(), 1, None, ws.data_ptr(), ws.numel(), _lib.ALGO_AUTO, _lib.stream())
    torch.cuda.synchronize()
    oc, ou = O.hsq_search(x.reshape(-1, 16), big)
    assert np.array_equal(codes.cpu().numpy(), oc) and np.array_equal(u.cpu().numpy(), ou)
