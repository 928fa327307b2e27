"""Drop-in for the reference's `compressors` package (compressors/__init__.py:1-8):
same class names, constructor arguments and compress()/decompress() signatures,
backed by the sm_100a kernels of libgqb200.so.  CUDA only."""
from .identical_compressor import IdenticalCompressor
from .qsgd_compressor import QSGDCompressor
from .probabilistic_scalar_compressor import ProbabilisticScalarCompressor
from .probabilistic_vector_compressor import ProbabilisticVectorCompressor
from .nearest_neighbor_compressor import NearestNeighborCompressor
from .residual_compressor import ResidualCompressor
from .signsgd_compressor import SignSGDCompressor
from .topk_sparsification_compressor import TopKSparsificationCompressor

__all__ = [
    "IdenticalCompressor", "QSGDCompressor", "ProbabilisticVectorCompressor",
    "NearestNeighborCompressor", "ResidualCompressor", "SignSGDCompressor",
    "TopKSparsificationCompressor", "ProbabilisticScalarCompressor",
]
