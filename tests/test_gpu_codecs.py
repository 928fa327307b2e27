"""GPU parity tests: QSGD/TernGrad, SignSGD, top-k, PVC/residual."""
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


@pytest.mark.parametrize("name", golden_names("qsgd_") + golden_names("terngrad_"))
def test_qsgd_golden(name):
    g = golden(name)
    shape = tuple(int(x) for x in g["shape"])
    size = int(np.prod(shape))
    a = make_args(c_dim=int(g["c_dim"]), n_bit=int(g["n_bit"]), random=bool(g["random"]))
    c = gq_b200.QSGDCompressor(size, torch.Size(shape), a)
    assert c.dim == int(g["dim"])
    x = _t(gen_input(int(g["seed"]), size, str(g["kind"]))).view(shape)
    norm, signs, l = c.compress(x, uniforms=torch_uniform_stream(int(g["seed"]), size))
    assert norm.shape == (c.M, 1) and signs.dtype == torch.bool and l.dtype == torch.int32
    assert signs.shape == shape and l.shape == shape
    assert np.array_equal(norm.cpu().numpy().reshape(-1), g["norm"])
    assert np.array_equal(signs.cpu().numpy().reshape(-1).astype(np.uint8), g["signs"])
    assert np.array_equal(l.cpu().numpy().reshape(-1), g["l"])      # incl. INT32_MIN on 0/0 chunks
    dec = c.decompress([norm, signs, l])
    assert np.array_equal(dec.cpu().numpy().reshape(-1), g["decoded"])


@pytest.mark.parametrize("c_dim,n_bit", [(128, 2), (0, 1), (128, 4), (128, 8), (64, 6)])
def test_qsgd_packed_wire_roundtrip(c_dim, n_bit):
    """packed record -> decode-reduce equals the unpacked reference-dtype path, per user and averaged."""
    from gq_b200.quantizers.fused import FusedPlan
    shapes = [(64, 128), (100,), (32, 64), (40, 3, 3, 4)]   # 1440 -> dim 288 (c_dim 128) / 96 (64)
    U = 3
    a = make_args(c_dim=c_dim, n_bit=n_bit, num_users=U)
    plan = FusedPlan(gq_b200.QSGDCompressor, shapes, a, torch.device(DEV), U)
    sizes = [int(np.prod(s)) for s in shapes]
    codecs = [O.QSGD(n, s, c_dim, n_bit, True) if n > 1000 else O.Identity() for n, s in zip(sizes, shapes)]
    n_draws = sum(n for n in sizes if n > 1000)
    decs = []
    for u in range(U):
        xs = [gen_input(300 + 10 * u + i, n).reshape(s) for i, (n, s) in enumerate(zip(sizes, shapes))]
        stream = torch_uniform_stream(70 + u, n_draws)
        parts, used = plan.split_uniform_stream(stream)
        plan.gather([_t(x) for x in xs])
        plan.encode(u, uniforms=parts)
        s = O.UniformStream(stream)
        decs.append([c.decompress(c.compress(x, s)) for c, x in zip(codecs, xs)])
    out = plan.decode(mean=True)
    for i in range(len(shapes)):
        ref = O.ps_mean(np.stack([decs[u][i].reshape(-1) for u in range(U)]))
        assert np.array_equal(plan.view(i, out).cpu().numpy().reshape(-1), ref), i


@pytest.mark.parametrize("c_dim,n_bit,shapes", [
    (8, 2, [(64, 128), (100,), (32, 64), (3, 2048)]),          # 8 lanes per chunk
    (32, 4, [(64, 128), (100,), (32, 64), (3, 2048)]),
    (512, 2, [(64, 128), (100,), (32, 64), (3, 2048)]),        # 4 float4 per lane
    (2048, 6, [(64, 128), (100,), (32, 64), (3, 2048)]),       # 16 float4 per lane
    (4096, 2, [(64, 128), (100,), (2, 4096)]),                 # beyond the register-resident kernel: generic path
    (50, 2, [(30, 100), (7,), (25, 80)]),                      # chunk dim not a multiple of 4: generic path
    (100, 8, [(30, 100), (7,), (25, 80)]),                     # 25 float4 per chunk (one lane group, partly idle)
])
def test_qsgd_every_chunk_layout_against_the_oracle(c_dim, n_bit, shapes):
    """Every instantiation of the one-launch QSGD encode (lanes per chunk, float4 per lane) and the generic
    fallbacks, through the packed record and the fused decode, bit for bit against the oracle."""
    from gq_b200.quantizers.fused import FusedPlan
    U = 2
    a = make_args(c_dim=c_dim, n_bit=n_bit, num_users=U)
    plan = FusedPlan(gq_b200.QSGDCompressor, shapes, a, torch.device(DEV), U)
    sizes = [int(np.prod(s)) for s in shapes]
    codecs = [O.QSGD(n, s, c_dim, n_bit, True) if n > 1000 else O.Identity() for n, s in zip(sizes, shapes)]
    n_draws = sum(n for n in sizes if n > 1000)
    decs = []
    for u in range(U):
        xs = [gen_input(900 + 10 * u + i, n).reshape(s) for i, (n, s) in enumerate(zip(sizes, shapes))]
        xs[0].reshape(-1)[: max(c_dim, 8)] = 0.0                      # an all-zero chunk: 0 / 0 levels
        stream = torch_uniform_stream(170 + u, n_draws)
        parts, used = plan.split_uniform_stream(stream)
        plan.gather([_t(x) for x in xs])
        plan.encode(u, uniforms=parts)
        st = O.UniformStream(stream)
        decs.append([c.decompress(c.compress(x, st)) for c, x in zip(codecs, xs)])
    out = plan.decode(mean=True)
    for i in range(len(shapes)):
        ref = O.ps_mean(np.stack([decs[u][i].reshape(-1) for u in range(U)]))
        assert np.array_equal(plan.view(i, out).cpu().numpy().reshape(-1), ref), i


def test_qsgd_max_element_decodes_exactly():
    a = make_args(c_dim=128, n_bit=2)
    x = gen_input(2, 128 * 50)
    c = gq_b200.QSGDCompressor(x.size, torch.Size((x.size,)), a)
    dec = c.decompress(c.compress(_t(x))).cpu().numpy().reshape(-1, 128)
    xm = x.reshape(-1, 128)
    j = np.abs(xm).argmax(1)
    assert np.array_equal(dec[np.arange(50), j], xm[np.arange(50), j])


def test_sign_and_topk_golden():
    g = golden("sign_topk")
    shape = tuple(int(x) for x in g["shape"])
    size = int(np.prod(shape))
    x = gen_input(int(g["seed"]), size)
    x[::7] = 0.0
    xt = _t(x).view(shape)
    s = gq_b200.SignSGDCompressor(size, torch.Size(shape), make_args())
    sg = s.compress(xt)
    assert np.array_equal(sg.cpu().numpy().reshape(-1).astype(np.int8), g["sign"])
    assert s.decompress(sg) is sg
    assert torch.equal(s.compress(sg), sg)                      # idempotent
    for cr in (100, 256, 3):
        c = gq_b200.TopKSparsificationCompressor(size, torch.Size(shape), make_args(cr=cr))
        sig = c.compress(xt)
        assert sig.shape == (1, size)
        out = c.decompress(sig)
        assert tuple(out.shape) == shape
        o = out.cpu().numpy().reshape(-1)
        assert np.array_equal(np.flatnonzero(o).astype(np.int32), g["topk_idx_cr%d" % cr])
        ref = O.topk(x, size // cr)
        assert np.array_equal(o, ref) and np.array_equal(np.signbit(o), np.signbit(ref))
        assert torch.equal(c.compress(out), sig)                # idempotent


@pytest.mark.parametrize("n,k", [(1, 1), (1000, 0), (1025, 1025), (5000, 17), (70001, 700), (1 << 20, 4096)])
def test_topk_sizes_and_ties(n, k):
    rs = np.random.RandomState(n + k)
    x = rs.standard_normal(n).astype(np.float32)
    x[rs.rand(n) < 0.5] = 0.0                       # many exact ties at zero
    x[:: max(n // 50, 1)] = 0.75                    # ties above the cut too
    ref = O.topk(x, k)
    seg = torch.tensor([0, n], dtype=torch.int64, device=DEV)
    kt = torch.tensor([k], dtype=torch.int64, device=DEV)
    kp = torch.zeros(1, dtype=torch.int64, device=DEV)
    ws = torch.empty(_lib.value("gq_topk_workspace_bytes", n, 1), dtype=torch.uint8, device=DEV)
    xt = _t(x)
    dense = torch.empty(n, device=DEV)
    idx = torch.full((max(k, 1),), -1, dtype=torch.int32, device=DEV)
    val = torch.zeros(max(k, 1), device=DEV)
    _lib.call("gq_topk_select", xt.data_ptr(), n, seg.data_ptr(), kt.data_ptr(), kp.data_ptr(), 1,
              dense.data_ptr(), idx.data_ptr(), val.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream())
    o = dense.cpu().numpy()
    assert np.array_equal(o, ref)
    want = np.sort(np.argsort(-np.abs(x), kind="stable")[:k]).astype(np.int32)
    if k:
        assert np.array_equal(idx.cpu().numpy()[:k], want)      # ascending index, lowest index wins ties
        assert np.array_equal(val.cpu().numpy()[:k], x[want])


def test_topk_segmented_and_sparse_mean():
    from gq_b200.quantizers.fused import FusedPlan
    shapes = [(64, 128), (100,), (32, 64), (40, 3, 3, 3), (3000,)]
    U = 4
    a = make_args(cr=100, num_users=U)
    plan = FusedPlan(gq_b200.TopKSparsificationCompressor, shapes, a, torch.device(DEV), U)
    sizes = [int(np.prod(s)) for s in shapes]
    decs = []
    for u in range(U):
        xs = [gen_input(500 + 10 * u + i, n).reshape(s) for i, (n, s) in enumerate(zip(sizes, shapes))]
        plan.gather([_t(x) for x in xs])
        plan.encode(u)
        decs.append([O.topk(x, x.size // 100).reshape(x.shape) if x.size > 1000 else x for x in xs])
    out = plan.decode(mean=True)
    for i in range(len(shapes)):
        ref = O.ps_mean(np.stack([decs[u][i].reshape(-1) for u in range(U)]))
        assert np.array_equal(plan.view(i, out).cpu().numpy().reshape(-1), ref), i


def test_sign_packed_mean():
    from gq_b200.quantizers.fused import FusedPlan
    shapes = [(64, 128), (100,), (1031,)]
    U = 5
    plan = FusedPlan(gq_b200.SignSGDCompressor, shapes, make_args(num_users=U), torch.device(DEV), U)
    sizes = [int(np.prod(s)) for s in shapes]
    decs = []
    for u in range(U):
        xs = [gen_input(600 + 10 * u + i, n).reshape(s) for i, (n, s) in enumerate(zip(sizes, shapes))]
        xs[0][0, :5] = 0.0
        plan.gather([_t(x) for x in xs])
        plan.encode(u)
        decs.append([O.sign(x).reshape(x.shape) if x.size > 1000 else x for x in xs])
    out = plan.decode(mean=True)
    for i in range(len(shapes)):
        ref = O.ps_mean(np.stack([decs[u][i].reshape(-1) for u in range(U)]))
        assert np.array_equal(plan.view(i, out).cpu().numpy().reshape(-1), ref), i


@pytest.mark.parametrize("U", [1, 2, 3, 8])
def test_sign_base3_wire(U):
    """args.sign_wire = "t5" (five ternary digits per byte): the wire bytes equal the oracle's
    packing of the reference's sign() output, the record shrinks to 1.6 bits per element, and the fused
    decode gives the 2-bit wire's values bit for bit -- ragged sizes (not multiples of 20), exact zeros."""
    from gq_b200.quantizers.fused import FusedPlan
    shapes = [(64, 128), (100,), (1031,), (7, 643), (20 * 640 + 13,)]
    a5, a2 = make_args(num_users=U, sign_wire="t5"), make_args(num_users=U)
    plan = FusedPlan(gq_b200.SignSGDCompressor, shapes, a5, torch.device(DEV), U)
    plan2 = FusedPlan(gq_b200.SignSGDCompressor, shapes, a2, torch.device(DEV), U)
    sizes = [int(np.prod(s)) for s in shapes]
    g = plan.groups[0]
    assert g.kind == "sign" and g.t5 and g.wire_bytes == (g.n + 19) // 20 * 4
    assert plan.wire_bytes() < plan2.wire_bytes() and g.wire_bytes * 5 <= (g.n + 19) // 20 * 20
    decs = []
    for u in range(U):
        xs = [gen_input(650 + 10 * u + i, n).reshape(s) for i, (n, s) in enumerate(zip(sizes, shapes))]
        xs[0][0, :7] = 0.0
        xs[3][2, 100:140] = -0.0
        for p_ in (plan, plan2):
            p_.gather([_t(x) for x in xs])
            p_.encode(u)
        # the wire itself: group order = tensors with more than 1000 elements, in plan order
        flat = np.concatenate([O.sign(xs[i]).reshape(-1) for i in g.tensors])
        wire = plan.records[u][g.packed_off:g.packed_off + g.wire_bytes].cpu().numpy()
        assert np.array_equal(wire, O.sign_pack_t5(flat)), u
        assert np.array_equal(O.sign_unpack_t5(wire, g.n), flat)
        decs.append([O.sign(x).reshape(x.shape) if x.size > 1000 else x for x in xs])
    for mean in (True, False):
        out, out2 = plan.decode(mean=mean), plan2.decode(mean=mean)
        assert torch.equal(out, out2)
        for i in range(len(shapes)):
            st = np.stack([decs[u][i].reshape(-1) for u in range(U)])
            if mean:
                ref = O.ps_mean(st)
            else:           # plain sum in user order, fp32
                ref = st[0].copy()
                for u in range(1, U):
                    ref = (ref + st[u]).astype(np.float32)
            assert np.array_equal(plan.view(i, out).cpu().numpy().reshape(-1), ref), (mean, i)
    # decode-accumulate (ring hop) and decode-subtract (error feedback) agree with the 2-bit wire too
    base = torch.randn_like(plan.arena)
    for acc in (1, 2):
        o5, o2 = base.clone(), base.clone()
        plan.decode(first_user=0, n_users=1, mean=False, accumulate=acc, out=o5)
        plan2.decode(first_user=0, n_users=1, mean=False, accumulate=acc, out=o2)
        assert torch.equal(o5, o2), acc


@pytest.mark.parametrize("d,k_bit", [(16, 8), (8, 8), (32, 8)])
def test_pvc_and_residual_vs_oracle(d, k_bit):
    """PVC implements the reference's intended algorithm (parity unpinned); the CUDA kernel
    must equal the oracle's restatement bit for bit, and be unbiased."""
    K = 2 ** k_bit
    n_chunks = 3000
    size = n_chunks * d
    x = gen_input(900 + d, size)
    a = make_args(c_dim=d, k_bit=k_bit, n_bit=6)
    cb = codebook(d, K)
    r1 = torch_uniform_stream(1, n_chunks)
    r2 = torch_uniform_stream(2, n_chunks)
    pvc = gq_b200.ProbabilisticVectorCompressor(size, torch.Size((size,)), a)
    assert np.array_equal(pvc.c_dagger.cpu().numpy(), np.linalg.pinv(cb.T).astype(np.float32))
    sig = pvc.compress(_t(x), uniforms=r1, norm_uniforms=r2)
    op = O.PVC(size, (size,), cb, 6, True)
    osig = op.compress(x, O.UniformStream(np.concatenate([r1, r2])))
    assert np.array_equal(sig[1].cpu().numpy().astype(np.int32), osig[1])
    assert np.float32(sig[0][0].item()) == osig[0][0] and np.float32(sig[0][1].item()) == osig[0][1]
    assert np.array_equal(sig[0][2].cpu().numpy(), osig[0][2])
    assert np.array_equal(pvc.decompress(sig).cpu().numpy(), op.decompress(osig))
    # residual: stage 1 HSQ + stage 2 PVC
    res = gq_b200.ResidualCompressor(size, torch.Size((size,)), a)
    r0 = torch_uniform_stream(3, n_chunks)
    sigs = res.compress(_t(x), uniforms=[dict(uniforms=r0), dict(uniforms=r1, norm_uniforms=r2)])
    orc = O.Residual(size, (size,), cb, 6, True)
    osigs = orc.compress(x, O.UniformStream(np.concatenate([r0, r1, r2])))
    assert np.array_equal(sigs[0][1].cpu().numpy().astype(np.int32), osigs[0][1])
    assert np.array_equal(sigs[1][1].cpu().numpy().astype(np.int32), osigs[1][1])
    assert np.array_equal(res.decompress(sigs).cpu().numpy(), orc.decompress(osigs))


@pytest.mark.parametrize("name", golden_names("pvc_") + golden_names("residual_"))
def test_pvc_and_residual_vs_reference_golden(name):
    """The reference's own ProbabilisticVectorCompressor / ResidualCompressor outputs
    (tests/golden/make_golden.py:pvc_case), codes / levels / lb-ub / decoded bit for bit."""
    g = golden(name)
    n, d, k_bit, n_bit = int(g["n_chunks"]), int(g["d"]), int(g["k_bit"]), int(g["n_bit"])
    size = n * d
    x = gen_input(int(g["seed"]), size)
    a = make_args(c_dim=d, k_bit=k_bit, n_bit=n_bit)
    draws = g["draws"]
    cb = torch.from_numpy(g["codewords"]).to(DEV)
    dag = torch.from_numpy(g["dagger"]).to(DEV)

    def check(sig, pre):
        assert np.array_equal(sig[1].cpu().numpy().astype(np.int32), g[pre + "codes"])
        if n_bit != 32:
            assert np.float32(sig[0][0].item()) == g[pre + "lb"] and np.float32(sig[0][1].item()) == g[pre + "ub"]
            assert np.array_equal(sig[0][2].cpu().numpy(), g[pre + "l"])
        else:
            assert np.array_equal(sig[0].cpu().numpy(), g[pre + "u"])

    if int(g["residual"]):
        c = gq_b200.ResidualCompressor(size, torch.Size((n, d)), a)
        assert np.array_equal(c.compressors[0].codewords.cpu().numpy(), g["codewords"])
        c.compressors[1].codewords, c.compressors[1].c_dagger = cb, dag
        sig = c.compress(_t(x).view(n, d), uniforms=[dict(uniforms=draws[:n]),
                                                      dict(uniforms=draws[n:2 * n], norm_uniforms=draws[2 * n:])])
        check(sig[0], "s1_")
        check(sig[1], "s2_")
    else:
        c = gq_b200.ProbabilisticVectorCompressor(size, torch.Size((n, d)), a)
        c.codewords, c.c_dagger = cb, dag       # K == d: the reference drew a random orthogonal basis
        sig = c.compress(_t(x).view(n, d), uniforms=draws[:n], norm_uniforms=draws[n:] if n_bit != 32 else None)
        check(sig, "")
    assert np.array_equal(c.decompress(sig).cpu().numpy().reshape(-1), g["decoded"])


def test_pvc_is_unbiased():
    d, size = 16, 16 * 64
    a = make_args(c_dim=d, k_bit=8, n_bit=32)
    x = gen_input(4, size)
    pvc = gq_b200.ProbabilisticVectorCompressor(size, torch.Size((size,)), a)
    xt = _t(x)
    acc = torch.zeros(size, device=DEV, dtype=torch.float64)
    T = 4000
    for _ in range(T):
        acc += pvc.decompress(pvc.compress(xt)).double()
    err = (acc / T - xt.double()).abs().max().item()
    assert err < 0.25 * np.abs(x).max(), err
