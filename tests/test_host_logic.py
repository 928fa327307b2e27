"""CPU: host-side logic -- dim rule, fvecs I/O, fused-plan layout, drop-in names."""
import os
import sys

import numpy as np
import torch

import gq_b200
from gq_b200.compressors._common import chunk_dim, load_codebook
from gq_b200.quantizers.fused import FusedPlan
from gq_b200.utils import vecs_io
from oracle import gq_oracle as O
from util import FCN_SHAPES, codebook, make_args, resnet50_shapes


def test_chunk_dim_rule():
    assert chunk_dim(1728, 16) == 16          # first ResNet conv divides by 16
    assert chunk_dim(1080, 16) == 24          # 16 -> 24
    assert chunk_dim(1728, 128) == 192        # QSGD d=128 -> 192
    assert chunk_dim(200704, 16) == 16
    assert chunk_dim(10, 16) == 10            # smaller than c_dim -> whole tensor
    assert chunk_dim(4096, 0) == 4096         # c_dim == 0 -> whole tensor (TernGrad)
    for size in (1728, 2560, 200704, 1000, 37, 65536):
        for c in (0, 8, 16, 32, 128):
            assert chunk_dim(size, c) == O.chunk_dim(size, c)


def test_codebook_rows_are_unit_norm_and_match_oracle():
    for d, K in ((8, 256), (16, 256), (32, 256), (16, 4096)):
        cb = load_codebook(d, K)
        assert cb.shape == (K, d) and cb.dtype == np.float32
        assert np.allclose(np.linalg.norm(cb.astype(np.float64), axis=1), 1.0, atol=1e-6)
        assert np.array_equal(cb, codebook(d, K))


def test_fvecs_roundtrip(tmp_path):
    a = np.random.RandomState(0).standard_normal((7, 5)).astype(np.float32)
    p = str(tmp_path / "x.fvecs")
    vecs_io.fvecs_writer(p, a)
    assert os.path.getsize(p) == 7 * 6 * 4
    assert np.array_equal(vecs_io.fvecs_read(p), a)
    assert np.array_equal(np.asarray(vecs_io.mmap_fvecs(p)), a)
    q = str(tmp_path / "x.ivecs")
    vecs_io.ivecs_writer(q, np.arange(12).reshape(3, 4))
    assert np.array_equal(vecs_io.ivecs_read(q), np.arange(12).reshape(3, 4))


def test_fused_plan_layout_resnet50_hsq():
    plan = FusedPlan(gq_b200.NearestNeighborCompressor, resnet50_shapes(), make_args(), torch.device("cpu"), 2)
    assert plan.total_elems() == 23520842
    assert plan.compressed_elems() == 23498432
    kinds = [g.kind for g in plan.groups]
    assert kinds == ["hsq", "identity"]
    g16, gid = plan.groups
    assert g16.key == (16, 256)
    assert g16.n_seg == 76 and g16.n_chunks == 23498432 // 16 == 1468652
    assert gid.n == 22410
    # every tensor has its own disjoint slot, groups are 256-byte aligned
    spans = sorted((plan.tensor_off[i], plan.tensor_off[i] + plan.sizes[i]) for i in range(len(plan.sizes)))
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 <= b0
    for g in plan.groups:
        assert (g.arena_off * 4) % 256 == 0
    # wire format: 2 bytes per chunk + lb/ub + identity fp32
    assert plan.wire_bytes() == 2 * g16.n_chunks + 8 * 76 + 4 * 22410
    assert plan.record_bytes % 256 == 0 and plan.records.shape == (2, plan.record_bytes)
    v = plan.view(0)
    assert v.shape == (64, 3, 3, 3)


def test_fused_plan_layout_other_codecs():
    a = make_args(c_dim=128, n_bit=2)
    plan = FusedPlan(gq_b200.QSGDCompressor, resnet50_shapes(), a, torch.device("cpu"), 1)
    assert [g.key for g in plan.groups if g.kind == "qsgd"] == [192, 128]
    tern = FusedPlan(gq_b200.QSGDCompressor, FCN_SHAPES, make_args(c_dim=0, n_bit=1), torch.device("cpu"), 1)
    g = tern.groups[0]
    assert g.key == "tensor" and g.n_chunks == 2 and g.bits == 4
    topk = FusedPlan(gq_b200.TopKSparsificationCompressor, FCN_SHAPES, make_args(cr=100), torch.device("cpu"), 1)
    assert topk.groups[0].k_total == 200704 // 100 + 2560 // 100
    sign = FusedPlan(gq_b200.SignSGDCompressor, FCN_SHAPES, make_args(), torch.device("cpu"), 1)
    assert sign.wire_bytes() == (203264 + 3) // 4 + 4 * 266


def test_split_uniform_stream_follows_reference_call_order():
    plan = FusedPlan(gq_b200.NearestNeighborCompressor, [(45, 24), (10,), (64, 64), (32, 64)],
                     make_args(), torch.device("cpu"), 1)
    n24, n16a, n16b = 1080 // 24, 4096 // 16, 2048 // 16
    stream = np.arange(n24 + n16a + n16b, dtype=np.float32)
    parts, used = plan.split_uniform_stream(stream)
    assert used == stream.size
    g24 = [g for g in plan.groups if g.kind == "hsq" and g.dim == 24][0]
    g16 = [g for g in plan.groups if g.kind == "hsq" and g.dim == 16][0]
    assert parts[id(g24)].tolist() == list(range(n24))
    assert parts[id(g16)].tolist() == list(range(n24, n24 + n16a + n16b))


def test_install_dropin_registers_reference_module_names():
    saved = {k: sys.modules.get(k) for k in ("compressors", "quantizers", "utils", "utils.vecs_io", "utils.vec_np")}
    try:
        gq_b200.install_dropin()
        import compressors
        import quantizers
        from utils.vecs_io import fvecs_read  # noqa: F401
        assert compressors.NearestNeighborCompressor is gq_b200.NearestNeighborCompressor
        assert quantizers.Quantizer is gq_b200.Quantizer
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_plan_locates_arena_shaped_gradients_and_counts_launches():
    a = make_args()
    plan = FusedPlan(gq_b200.NearestNeighborCompressor, FCN_SHAPES, a, torch.device("cpu"), 2)
    buf = torch.zeros(plan.arena_elems + 64)
    off = (-buf.data_ptr() // 4) % 64            # a 256-byte aligned start inside buf
    flat = buf[off:off + plan.arena_elems]
    views = plan.views(flat)
    assert plan.locate(views) == flat.data_ptr()                 # laid out like the arena: read in place
    moved = list(views)
    moved[1] = torch.zeros_like(views[1])                        # one tensor lives elsewhere: copy path
    assert plan.locate(moved) is None
    assert plan.locate(views[:-1]) is None
    # ResNet-50 / HSQ d=16 K=256 n=6: ONE encode launch (key reset, identity copy, search, fused norm
    # quantization) and one decode kernel that also reduces the identity tensors
    big = FusedPlan(gq_b200.NearestNeighborCompressor, resnet50_shapes(), a, torch.device("cpu"), 2)
    assert big.launches_per_encode() == 1 and big.launches_per_decode(2) == 1
    assert big.supports_fused_delivery()
    from gq_b200 import _lib
    exact = FusedPlan(gq_b200.NearestNeighborCompressor, resnet50_shapes(), make_args(hsq_algo=_lib.ALGO_EXACT),
                      torch.device("cpu"), 2)
    assert exact.launches_per_encode() == 3                      # init + search + quantize
    d8 = FusedPlan(gq_b200.NearestNeighborCompressor, resnet50_shapes(), make_args(c_dim=8), torch.device("cpu"), 2)
    # d = 8 (and 32) run the one-launch tcgen05 encode too; their decode kernel does not carry the identity reduce
    assert d8.launches_per_encode() == 1 and d8.launches_per_decode(2) == 2
    assert d8.supports_fused_delivery() and not exact.supports_fused_delivery()


def test_peer_records_row_and_address_arithmetic():
    from gq_b200.quantizers.p2p import PeerRecords
    pr = object.__new__(PeerRecords)
    pr.record_bytes, pr.rank, pr.world, pr.step = 4096, 2, 4, 0
    pr.base = [0x1000000 * (r + 1) for r in range(4)]
    assert pr.parity == 0 and pr.row() == 2 and pr.row(0) == 0
    assert pr._addr(1, 3) == pr.base[1] + 3 * 4096               # user 3's row inside rank 1's block
    offs = list(pr.user_offsets())
    assert offs[0] == 0 and offs[3] == (pr.base[3] + 3 * 4096) - pr.base[0]
    pr.step = 1                                                  # odd step: second half of the block
    assert pr.parity == 1 and pr.row() == 4 + 2 and pr.user0_record_ptr() == pr.base[0] + 4 * 4096


def test_codebook_generator_small(tmp_path):
    """gq_b200.codebook_generator: the reference's train_codebook / generate surface
    (codebook_generator.py:14-31), files readable by fvecs_read and by load_codebook."""
    from gq_b200 import codebook_generator as G
    from gq_b200.utils.vecs_io import fvecs_read
    cb = G.train_codebook(8, 32, train_size=4000, iter=5, device=torch.device("cpu"))
    assert cb.shape == (32, 8) and cb.dtype == np.float32 and np.isfinite(cb).all()
    again = G.train_codebook(8, 32, train_size=4000, iter=5, device=torch.device("cpu"))
    assert np.array_equal(cb, again)                                  # deterministic for a seed
    rnd = np.random.RandomState(0).standard_normal((32, 8)).astype(np.float32)
    assert G.quantisation_error(cb, n=4000) < G.quantisation_error(rnd, n=4000)   # better than random directions
    out = G.generate(dims=[4], Ks=[8], out_dir=str(tmp_path), train_size=500, iter=3)
    assert len(out) == 1 and out[0].endswith("angular_dim_4_Ks_8.fvecs")
    assert fvecs_read(out[0]).shape == (8, 4)
    assert G.generate(dims=[4], Ks=[8], out_dir=str(tmp_path), train_size=500, iter=3) == []   # kept, not appended to


def test_ring_stages_cut_the_resnet50_plan_at_aligned_tensor_boundaries():
    """make_parts on the headline workload: the first convolution has 108 chunks, so the plan places it
    after the tensors whose chunk count is a multiple of 16 and four real stages come out."""
    plan = FusedPlan(gq_b200.NearestNeighborCompressor, resnet50_shapes(), make_args(), torch.device("cpu"), 2)
    assert plan.make_parts(4) == 4
    g = plan.groups[0]
    covered = 0
    for p in range(4):
        ta, tb, ca, cb = plan._part_range(g, p)
        assert ca % 16 == 0 and ca == covered and tb > ta
        assert 0.15 * g.n_chunks < cb - ca < 0.40 * g.n_chunks        # about a quarter each
        covered = cb
    assert covered == g.n_chunks
    # the ranges of all stages tile the codes / levels / lb-ub sections of the record exactly once
    seen = np.zeros(plan.record_bytes, np.int32)
    for p in range(4):
        for b, e in plan.part_byte_ranges(p):
            seen[b:e] += 1
    for off, n in ((g.codes_off, g.n_chunks), (g.l_off, g.n_chunks), (g.lbub_off, 8 * g.n_seg)):
        assert (seen[off:off + n] == 1).all()


def test_sign_wire_containers_in_the_plan_layout():
    """SignSGD record: 2 bits per element by default, base 3 (five elements per byte, whole 32-bit words) with
    args.sign_wire = "t5" (INTEGRATION.md section 3); everything else of the layout is unchanged."""
    shapes = resnet50_shapes()
    p2 = FusedPlan(gq_b200.SignSGDCompressor, shapes, make_args(), torch.device("cpu"), 1)
    p5 = FusedPlan(gq_b200.SignSGDCompressor, shapes, make_args(sign_wire="t5"), torch.device("cpu"), 1)
    g2, g5 = p2.groups[0], p5.groups[0]
    assert g2.kind == g5.kind == "sign" and g2.n == g5.n == 23498432
    assert not g2.t5 and g2.wire_bytes == g2.n // 4
    assert g5.t5 and g5.wire_bytes == (g5.n + 19) // 20 * 4
    assert p5.wire_bytes() == p2.wire_bytes() - g2.wire_bytes + g5.wire_bytes
    assert 0.79 < g5.wire_bytes / g2.wire_bytes < 0.81            # 1.6 instead of 2 bits per element
    assert p5.record_bytes % 256 == 0 and p5.tensor_off == p2.tensor_off


def test_identity_tensors_ride_in_the_largest_group_of_every_codec():
    """The copy / reduction of the tensors with <= 1000 elements is attached to the largest compressed group
    (its kernel has the threads to spare); more than 8 users: a launch of its own."""
    shapes = resnet50_shapes()
    for Comp, kw in ((gq_b200.NearestNeighborCompressor, {}), (gq_b200.QSGDCompressor, dict(c_dim=128, n_bit=2)),
                     (gq_b200.SignSGDCompressor, {}), (gq_b200.TopKSparsificationCompressor, dict(cr=100))):
        plan = FusedPlan(Comp, shapes, make_args(**kw), torch.device("cpu"), 2)
        ident, carrier = plan._rider_pair()
        assert ident is not None and ident.kind == "identity" and ident.n == 22410
        compressed = [g for g in plan.groups if g.kind != "identity"]
        assert carrier is max(compressed, key=lambda g: g.n)
        assert plan._rider_pair(n_users=9) == (None, None)
    qs = FusedPlan(gq_b200.QSGDCompressor, shapes, make_args(c_dim=128, n_bit=2), torch.device("cpu"), 2)
    assert [g.key for g in qs.groups if g.kind == "qsgd"] == [192, 128] and qs._rider_pair()[1].key == 128
    ide = FusedPlan(gq_b200.IdenticalCompressor, shapes, make_args(), torch.device("cpu"), 2)
    assert ide._rider_pair() == (None, None)
