"""CPU: the C-ABI library loads and exports every symbol include/gqb200.h declares."""
import ctypes
import os
import re

import gq_b200
from gq_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gqb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gq_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libgqb200.so does not export %s" % n


def test_python_binding_covers_header():
    assert set(declared_symbols()) == set(_lib.EXPORTS)


def test_version_and_host_side_queries():
    assert _lib.value("gq_abi_version") == 1
    assert _lib.value("gq_qsgd_wire_bits", 1) == 4 and _lib.value("gq_qsgd_wire_bits", 2) == 4
    assert _lib.value("gq_qsgd_wire_bits", 6) == 8 and _lib.value("gq_qsgd_wire_bits", 8) == 16
    assert _lib.value("gq_hsq_encode_workspace_bytes", 1000, 16, 256, 3) >= 24
    assert _lib.value("gq_topk_workspace_bytes", 100000, 4) > 4 * 2048 * 4


def test_reference_names_are_exported():
    for n in ("IdenticalCompressor", "QSGDCompressor", "ProbabilisticVectorCompressor",
              "NearestNeighborCompressor", "ResidualCompressor", "SignSGDCompressor",
              "TopKSparsificationCompressor", "Quantizer", "PSQuantizer", "RingQuantizer"):
        assert hasattr(gq_b200, n)


def test_no_cpu_path():
    import pytest
    import torch
    from util import make_args
    with pytest.raises(_lib.GQError):
        gq_b200.NearestNeighborCompressor(4096, (64, 64), make_args(no_cuda=True))
    if not torch.cuda.is_available():
        with pytest.raises(_lib.GQError):
            gq_b200.NearestNeighborCompressor(4096, (64, 64), make_args())
