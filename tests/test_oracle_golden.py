"""CPU: the oracle against every golden fixture produced by the live reference
(tests/golden/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

from oracle import gq_oracle as O
from util import FCN_SHAPES, codebook, gen_input, golden, golden_names, torch_uniform_stream


@pytest.mark.parametrize("name", golden_names("hsq_"))
def test_hsq_fixture(name):
    g = golden(name)
    shape = tuple(int(x) for x in g["shape"])
    size = int(np.prod(shape))
    dim, k_bit, n_bit = int(g["dim"]), int(g["k_bit"]), int(g["n_bit"])
    x = gen_input(int(g["seed"]), size, str(g["kind"]))
    assert O.chunk_dim(size, int(g["c_dim"])) == dim
    cb = codebook(dim, 2 ** k_bit)
    c = O.HSQ(size, shape, cb, n_bit, bool(g["random"]))
    # the recorded draws are the reference's torch.manual_seed(seed) stream
    assert np.array_equal(g["draws"], torch_uniform_stream(int(g["seed"]), g["draws"].size))
    sig = c.compress(x, O.UniformStream(g["draws"]))
    assert np.array_equal(sig[1], g["codes"])
    if n_bit != 32:
        assert sig[0][0] == g["lb"] and sig[0][1] == g["ub"]
        assert np.array_equal(sig[0][2], g["l"])
        assert sig[0][2].min() >= 0 and sig[0][2].max() <= 2 ** n_bit
    else:
        assert np.array_equal(sig[0], g["u"])
    assert np.array_equal(c.decompress(sig).reshape(-1), g["decoded"])


@pytest.mark.parametrize("name", golden_names("qsgd_") + golden_names("terngrad_"))
def test_qsgd_fixture(name):
    g = golden(name)
    shape = tuple(int(x) for x in g["shape"])
    size = int(np.prod(shape))
    x = gen_input(int(g["seed"]), size, str(g["kind"]))
    c = O.QSGD(size, shape, int(g["c_dim"]), int(g["n_bit"]), bool(g["random"]))
    assert c.dim == int(g["dim"])
    sig = c.compress(x, O.UniformStream(torch_uniform_stream(int(g["seed"]), size)))
    assert np.array_equal(sig[0], g["norm"])
    assert np.array_equal(sig[1], g["signs"])
    assert np.array_equal(sig[2], g["l"])
    assert np.array_equal(c.decompress(sig).reshape(-1), g["decoded"], equal_nan=True)


def test_sign_topk_fixture():
    g = golden("sign_topk")
    shape = tuple(int(x) for x in g["shape"])
    size = int(np.prod(shape))
    x = gen_input(int(g["seed"]), size)
    x[::7] = 0.0
    assert np.array_equal(O.sign(x).astype(np.int8), g["sign"])
    for cr in (100, 256, 3):
        out = O.topk(x, size // cr)
        assert np.array_equal(np.flatnonzero(out).astype(np.int32), g["topk_idx_cr%d" % cr])
        assert np.array_equal(out[out != 0], x[out != 0])


def _quantizer_replay(g):
    U, seed, iters = int(g["U"]), int(g["seed"]), int(g["iters"])
    shapes = [tuple(int(x) for x in g["shape%d" % i]) for i in range(int(g["n_tensors"]))]
    sizes = [int(np.prod(s)) for s in shapes]
    quant = str(g["quant"])
    codecs = []
    for s, n in zip(shapes, sizes):
        if n <= 1000:
            codecs.append(O.Identity())
        elif quant == "hsq":
            d = O.chunk_dim(n, int(g["c_dim"]))
            codecs.append(O.HSQ(n, s, codebook(d, 2 ** int(g["k_bit"])), int(g["n_bit"]), True))
        elif quant == "qsgd":
            codecs.append(O.QSGD(n, s, int(g["c_dim"]), int(g["n_bit"]), True))
        elif quant == "sign":
            codecs.append(O.Sign(n, s))
        else:
            codecs.append(O.TopK(n, s, int(g["cr"])))
    stream = O.UniformStream(torch_uniform_stream(seed, int(g["n_draws"])))
    errs = [[np.zeros(s, np.float32) for s in shapes] for _ in range(U)] if int(g["ef"]) else None
    two_phase = bool(int(g["two_phase"])) if "two_phase" in g else False
    serrs = [np.zeros(s, np.float32) for s in shapes] if (two_phase and int(g["ef"])) else None
    for it in range(iters):
        grads = [[gen_input(seed * 1000 + it * 100 + u * 10 + i, n).reshape(s)
                  for i, (n, s) in enumerate(zip(sizes, shapes))] for u in range(U)]
        if str(g["mode"]) == "ps":
            out = O.ps_step(codecs, grads, stream, errs, O.ps_scale(int(g["epoch"])), two_phase, serrs)
        else:
            out = O.ring_step(codecs, grads, stream, errs, O.ps_scale(int(g["epoch"])))
        for i in range(len(shapes)):
            ref = g["grad_it%d_t%d" % (it, i)]
            got = out[i].reshape(-1)
            # torch's stack().mean(0) does not sum tiny tensors in plain user order
            # (last-bit differences on the 10-element bias); everything else is exact.
            # Bar for averaged gradients (north_star): 1e-5 relative; we hold 1e-6.
            rel = np.abs(ref - got).max() / max(np.abs(ref).max(), 1e-30)
            assert rel <= 1e-6, (it, i, rel)
            if ref.size >= 256:
                assert np.array_equal(ref, got), (it, i, rel)


@pytest.mark.parametrize("name", golden_names("ps_") + golden_names("ring_"))
def test_quantizer_fixture(name):
    _quantizer_replay(golden(name))


def _check_sig(sig, g, pre, n_bit):
    assert np.array_equal(sig[1], g[pre + "codes"])
    if n_bit != 32:
        assert sig[0][0] == g[pre + "lb"] and sig[0][1] == g[pre + "ub"]
        assert np.array_equal(sig[0][2], g[pre + "l"])
    else:
        assert np.array_equal(sig[0], g[pre + "u"])


@pytest.mark.parametrize("name", golden_names("pvc_") + golden_names("residual_"))
def test_pvc_and_residual_fixture(name):
    """ProbabilisticVectorCompressor / ResidualCompressor outputs of the reference's own code
    (make_golden.py:_runnable_pvc documents the one line that had to be read, not run)."""
    g = golden(name)
    n, d, n_bit = int(g["n_chunks"]), int(g["d"]), int(g["n_bit"])
    x = gen_input(int(g["seed"]), n * d)
    cb = g["codewords"]
    assert np.array_equal(np.linalg.pinv(cb.T).astype(np.float32), g["dagger"])
    assert np.array_equal(g["draws"], torch_uniform_stream(int(g["seed"]), g["draws"].size))
    stream = O.UniformStream(g["draws"])
    if int(g["residual"]):
        c = O.Residual(n * d, (n, d), cb, n_bit, True)
        sig = c.compress(x, stream)
        _check_sig(sig[0], g, "s1_", n_bit)
        _check_sig(sig[1], g, "s2_", n_bit)
    else:
        c = O.PVC(n * d, (n, d), cb, n_bit, True)
        sig = c.compress(x, stream)
        _check_sig(sig, g, "", n_bit)
    assert stream.pos == g["draws"].size
    assert np.array_equal(c.decompress(sig).reshape(-1), g["decoded"])


def test_psc_levels_cover_2n_plus_1():
    """'n-bit' has 2^n + 1 levels: the maximum always rounds up to 2^n."""
    u = gen_input(3, 4096)
    r = torch_uniform_stream(3, 4096)
    lb, ub, l, used = O.psc_compress(u, 6, True, r)
    assert used and l.min() == 0 and l.max() == 64
    lb, ub, l, used = O.psc_compress(u, 6, False, None)
    assert l.max() == 63
    lb, ub, l, used = O.psc_compress(np.full(64, 0.25, np.float32), 6, True, r)
    assert not used and not l.any() and lb == ub


def test_hsq_codeword_input_recovers_itself():
    cb = codebook(16, 256)
    x = (cb[[3, 200, 17]] * np.array([[0.5], [-2.0], [1e-3]], np.float32)).astype(np.float32)
    codes, u = O.hsq_search(x, cb)
    assert list(codes) == [3, 200, 17]
    assert np.allclose(u, [0.5, -2.0, 1e-3], rtol=1e-6)


def test_sign_base3_wire_packing_roundtrip():
    """oracle.sign_pack_t5 / sign_unpack_t5 (the checker of gq_sign_encode_t5): five ternary digits per byte,
    sections padded to whole 32-bit words, every ternary vector survives the round trip."""
    rs = np.random.RandomState(3)
    for n in (1, 4, 5, 19, 20, 21, 640, 1031, 12813):
        sig = rs.randint(-1, 2, n).astype(np.float32)
        w = O.sign_pack_t5(sig)
        assert w.dtype == np.uint8 and w.size == (n + 19) // 20 * 4 and int(w.max()) <= 242
        assert np.array_equal(O.sign_unpack_t5(w, n), sig)
    # the digit code and the digit order
    assert O.sign_pack_t5(np.array([1, -1, 0, 0, 1], np.float32))[0] == 1 + 3 * 2 + 81 * 1
