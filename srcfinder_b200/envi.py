"""Minimal ENVI header / image I/O for the matched-filter products (no `spectral` dependency).

The reference reads its input and writes its products through spectral.io.envi
(cmf/robust_mf.py:16-17, :206-207, :261-263, :278-279).  Header *bytes* of that writer are not pinned by
anything in the reference, so parity is defined on parsed key/values: keys are
lower-cased, scalar values are strings, ``{a, b}`` values are lists of strings (``description`` stays a
string), exactly what the reference's header dict holds when it mutates it (:210-230).
"""
from __future__ import annotations

import os

import numpy as np

ENVI_TO_NUMPY = {"1": "u1", "2": "i2", "3": "i4", "4": "f4", "5": "f8", "12": "u2", "13": "u4",
                 "14": "i8", "15": "u8"}
NUMPY_TO_ENVI = {"uint8": "1", "int16": "2", "int32": "3", "float32": "4", "float64": "5",
                 "uint16": "12", "uint32": "13", "int64": "14", "uint64": "15"}
_LEADING = ["description", "samples", "lines", "bands", "header offset", "file type", "data type",
            "interleave", "byte order"]


def read_header(path):
    with open(path, "r") as fh:
        lines = fh.read().splitlines()
    if not lines or not lines[0].strip().upper().startswith("ENVI"):
        raise ValueError("%s: not an ENVI header" % path)
    meta, i = {}, 1
    while i < len(lines):
        line = lines[i]
        i += 1
        if "=" not in line:
            continue
        key, val = line.split("=", 1)
        key, val = key.strip().lower(), val.strip()
        if val.startswith("{"):
            while "}" not in val and i < len(lines):
                val += "\n" + lines[i]
                i += 1
            inner = val[val.index("{") + 1:val.rindex("}")]
            meta[key] = inner.strip() if key == "description" else \
                [tok.strip() for tok in inner.replace("\n", " ").split(",")]
        else:
            meta[key] = val
    return meta


def write_header(path, meta):
    keys = [k for k in _LEADING if k in meta] + [k for k in meta if k not in _LEADING]
    with open(path, "w") as fh:
        fh.write("ENVI\n")
        for k in keys:
            v = meta[k]
            if isinstance(v, (list, tuple)):
                fh.write("%s = { %s }\n" % (k, " , ".join(str(x) for x in v)))
            elif k == "description":
                fh.write("%s = {\n  %s}\n" % (k, v))
            else:
                fh.write("%s = %s\n" % (k, v))


def image_shape(meta):
    L, S, B = int(meta["lines"]), int(meta["samples"]), int(meta["bands"])
    il = str(meta.get("interleave", "bil")).lower()
    if il not in ("bil", "bip", "bsq"):
        raise ValueError("unknown interleave %r" % il)
    return {"bil": (L, B, S), "bip": (L, S, B), "bsq": (B, L, S)}[il]


def image_dtype(meta):
    dt = np.dtype(ENVI_TO_NUMPY[str(meta["data type"]).strip()])
    if str(meta.get("byte order", "0")).strip() == "1":
        dt = dt.newbyteorder(">")
    return dt


def open_memmap(data_path, meta=None, writable=False):
    """Memory map of an ENVI image in its native interleave ((L,B,S) for BIL, like :207-208)."""
    if meta is None:
        meta = read_header(data_path + ".hdr")
    return np.memmap(data_path, dtype=image_dtype(meta), mode="r+" if writable else "r",
                     offset=int(meta.get("header offset", 0)), shape=image_shape(meta))


def create_image(data_path, meta):
    """Write ``data_path + '.hdr'`` and allocate a zero-filled image; returns a writable memmap."""
    meta = dict(meta)
    # a product never inherits the input's header offset or byte order (spectral's create_image writes native
    # byte order at offset 0 whatever the metadata copied from the input file said, cmf/robust_mf.py:210-263)
    meta["header offset"] = 0
    meta["byte order"] = 0
    meta.setdefault("file type", "ENVI Standard")
    write_header(data_path + ".hdr", meta)
    nbytes = int(np.prod(image_shape(meta))) * image_dtype(meta).itemsize
    with open(data_path, "wb") as fh:
        fh.truncate(nbytes)
    return open_memmap(data_path, meta, writable=True)
