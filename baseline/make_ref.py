#!/usr/bin/env python
"""Stage the UNMODIFIED reference for the CPU baseline arm.

    python baseline/make_ref.py [/root/reference]

Copies compressors/, quantizers/, utils/ and the handful of learned codebooks the bench configurations
need from the reference checkout into baseline/_ref/ (git-ignored, NOT gpurun-ignored: it travels to
the GPU box with the snapshot, where /root/reference does not exist).  The reference has no
setup.py / pyproject.toml, so `pip install --target baseline/_ref /root/reference` has nothing to
build ("neither 'setup.py' nor 'pyproject.toml' found"): a plain copy of the needed files is the
install.  Nothing under baseline/_ref is ever imported by the product or committed.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
CODEBOOKS = [(8, 256), (16, 256), (24, 256), (32, 256), (8, 4096), (16, 4096), (32, 4096)]


def make(src="/root/reference"):
    if not os.path.isdir(src):
        return False
    for sub in ("compressors", "quantizers", "utils"):
        d = os.path.join(DST, sub)
        if os.path.isdir(d):
            shutil.rmtree(d)
        shutil.copytree(os.path.join(src, sub), d, ignore=shutil.ignore_patterns("__pycache__"))
    cb = os.path.join(DST, "codebooks", "learned_codebook")
    os.makedirs(cb, exist_ok=True)
    for d, k in CODEBOOKS:
        name = "angular_dim_%d_Ks_%d.fvecs" % (d, k)
        s = os.path.join(src, "codebooks", "learned_codebook", name)
        if os.path.exists(s):
            shutil.copyfile(s, os.path.join(cb, name))
    with open(os.path.join(DST, "README"), "w") as fh:
        fh.write("Unmodified files of xinyandai/gradient-quantization staged by baseline/make_ref.py "
                 "for the CPU baseline arm of bench.py.  Not product code; git-ignored.\n")
    return True


if __name__ == "__main__":
    ok = make(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print("baseline/_ref %s" % ("staged" if ok else "NOT staged (no reference checkout here)"))
