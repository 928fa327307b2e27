"""Row normalisation used when loading codebooks (reference utils/vec_np.py:4-10).

The arithmetic must stay numpy's: the codebook handed to the CUDA kernels has to
be bit-identical to the one the reference builds, so this uses the same two
numpy calls (np.linalg.norm, np.divide with a where-mask) and nothing else.
"""
import numpy as np


def normalize(vecs, order=None):
    """-> (norms [rows], vecs / norms[:, None]); all-zero rows stay zero."""
    lengths = np.linalg.norm(vecs, ord=order, axis=1)
    col = lengths.reshape(-1, 1)
    unit = np.zeros_like(vecs)
    np.divide(vecs, col, out=unit, where=(col != 0))
    return lengths, unit
