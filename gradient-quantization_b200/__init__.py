"""gradient-quantization_b200 -- B200 (sm_100a) implementation of the compression
hot path of xinyandai/gradient-quantization, behind the reference's own
`compressors` / `quantizers` API.

    import gq_b200                      # import shim at the repo root
    from gq_b200.compressors import NearestNeighborCompressor
    from gq_b200.quantizers import Quantizer

gq_b200.install_dropin() registers the sub-packages under the reference's
top-level names (`compressors`, `quantizers`, `utils`) so that the reference's
main.py (`from compressors import *`, `from quantizers import *`) runs on this
implementation unchanged.

All arithmetic runs in libgqb200.so (CUDA, C ABI in include/gqb200.h).  There is
no CPU fallback; importing this package on a machine without the built library
raises ImportError.
"""
import sys

from . import _lib

_lib.load()   # fail loudly, at import, if the CUDA library is missing

from . import compressors, quantizers, utils  # noqa: E402
from .compressors import *  # noqa: E402,F401,F403
from .quantizers import Quantizer, PSQuantizer, RingQuantizer  # noqa: E402,F401

__version__ = "0.1.0"


def install_dropin():
    """Make `import compressors`, `import quantizers`, `import utils` resolve here."""
    me = sys.modules[__name__]
    sys.modules["compressors"] = me.compressors
    sys.modules["quantizers"] = me.quantizers
    sys.modules["utils"] = me.utils
    for sub in ("vecs_io", "vec_np"):
        __import__(__name__ + ".utils." + sub)
        sys.modules["utils." + sub] = sys.modules[__name__ + ".utils." + sub]
    return me
