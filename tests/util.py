"""Shared helpers for the tests (inputs, fixtures, codebooks)."""
import glob
import os
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden")
CODEBOOKS = os.path.join(ROOT, "gradient-quantization_b200", "codebooks", "learned_codebook")


def gen_input(seed, n, kind="normal"):
    """Same recipe as tests/golden/make_golden.py:gen_input (frozen RandomState stream)."""
    rs = np.random.RandomState(seed)
    if kind == "normal":
        return (rs.standard_normal(n) * 0.01).astype(np.float32)
    if kind == "heavy":
        return (rs.standard_normal(n) * np.exp(rs.standard_normal(n) * 3.0) * 1e-3).astype(np.float32)
    if kind == "repeat16":
        return np.tile((rs.standard_normal(16) * 0.01).astype(np.float32), n // 16)
    if kind == "zeros_mixed":
        x = (rs.standard_normal(n) * 0.01).astype(np.float32)
        x[: n // 4] = 0.0
        return x
    raise ValueError(kind)


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def golden_names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def codebook(d, K):
    from oracle import gq_oracle as O
    return O.normalize(O.fvecs_read(os.path.join(CODEBOOKS, "angular_dim_%d_Ks_%d.fvecs" % (d, K))))[1]


def make_args(**kw):
    base = dict(c_dim=16, k_bit=8, n_bit=6, no_cuda=False, random=True, cr=256, ef=False,
                two_phase=False, mode="ps", scale="exp", num_users=8)
    base.update(kw)
    return SimpleNamespace(**base)


def torch_uniform_stream(seed, n):
    """The reference's CPU uniform stream: torch.manual_seed(seed); torch.rand(n)."""
    import torch
    g = torch.Generator()
    g.manual_seed(int(seed))
    return torch.rand(int(n), generator=g).numpy()


RESNET50_SHAPES = None


def resnet50_shapes():
    """Parameter shapes of the reference's CIFAR ResNet-50 (models/resnet.py:40-65,108-109):
    161 tensors, 23 520 842 parameters, in model.parameters() order."""
    global RESNET50_SHAPES
    if RESNET50_SHAPES is not None:
        return RESNET50_SHAPES
    shapes = [(64, 3, 3, 3), (64,), (64,)]
    inch = 64
    for planes, blocks, stride in ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)):
        for s in [stride] + [1] * (blocks - 1):
            shapes += [(planes, inch, 1, 1), (planes,), (planes,),
                       (planes, planes, 3, 3), (planes,), (planes,),
                       (planes * 4, planes, 1, 1), (planes * 4,), (planes * 4,)]
            if s != 1 or inch != planes * 4:
                shapes += [(planes * 4, inch, 1, 1), (planes * 4,), (planes * 4,)]
            inch = planes * 4
    shapes += [(10, 2048), (10,)]
    RESNET50_SHAPES = shapes
    return shapes


FCN_SHAPES = [(256, 784), (256,), (10, 256), (10,)]
