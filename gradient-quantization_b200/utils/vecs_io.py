"""Readers/writers for the .fvecs / .ivecs / .bvecs vector files the codebooks ship in.

Same function names and file format as the reference's utils/vecs_io.py:5-57 so
that code written against it keeps working.  Format: every row is a little-endian
int32 dimension d followed by d payload items (fp32, int32 or uint8).
"""
import numpy as np


def _rows(raw32, what):
    if raw32.size == 0:
        raise ValueError("%s: empty file" % what)
    d = int(raw32[0])
    if d <= 0 or raw32.size % (d + 1) != 0:
        raise ValueError("%s: not a vecs file (leading dim %d, %d words)" % (what, d, raw32.size))
    table = raw32.reshape(-1, d + 1)
    if not (table[:, 0] == d).all():
        raise ValueError("%s: rows with differing dimensions" % what)
    return table[:, 1:]


def ivecs_read(fname):
    """int32 matrix [rows, d]."""
    return np.ascontiguousarray(_rows(np.fromfile(fname, dtype="<i4"), fname))


def fvecs_read(fname):
    """fp32 matrix [rows, d] (the payload words reinterpreted, no conversion)."""
    return ivecs_read(fname).view(np.float32)


def mmap_fvecs(fname):
    """Memory-mapped fp32 view for files too large to load."""
    raw = np.memmap(fname, dtype="<i4", mode="r")
    d = int(raw[0])
    return raw.view(np.float32).reshape(-1, d + 1)[:, 1:]


def mmap_bvecs(fname):
    """Memory-mapped uint8 view: int32 d then d bytes per row."""
    raw = np.memmap(fname, dtype=np.uint8, mode="r")
    d = int(raw[:4].view("<i4")[0])
    return raw.reshape(-1, d + 4)[:, 4:]


def bvecs_read(filename):
    return mmap_bvecs(filename)


def _append(filename, vecs, dtype):
    mat = np.asarray(vecs, dtype=dtype)
    if mat.ndim != 2:
        raise ValueError("expected a 2-D array of row vectors")
    head = np.full((mat.shape[0], 1), mat.shape[1], dtype="<i4")
    with open(filename, "ab") as fh:  # append, like the reference writer
        np.concatenate([head.view(np.uint8).reshape(mat.shape[0], 4),
                        mat.view(np.uint8).reshape(mat.shape[0], -1)], axis=1).tofile(fh)


def fvecs_writer(filename, vecs):
    _append(filename, vecs, "<f4")


def ivecs_writer(filename, vecs):
    _append(filename, vecs, "<i4")
