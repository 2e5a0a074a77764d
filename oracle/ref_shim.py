"""TEST INFRASTRUCTURE -- not part of the product path.

Shim that lets the UNMODIFIED reference script ``/root/reference/cmf/robust_mf.py``
run in this container, so its outputs can pin the numpy restatement in
``oracle/cmf_oracle.py`` and generate the golden vectors under ``tests/golden/``.

Only ``oracle/make_golden.py`` (run once, in the build container, where
``/root/reference`` exists) uses this module.  Nothing here is imported by the
product package; ``/root/reference`` does not exist on the GPU box.

What has to be faked (reference lines are cmf/robust_mf.py):
  * ``spectral.io.envi``  -- ``open``/``create_image``/``dtype_to_envi``  (:16-17, :46-50)
  * module globals ``os``, ``sys``, ``np`` that the script uses but never binds
    (:183, :194, :293-295), and a list-returning ``map`` (:215-216, python-2 idiom)
  * the terminal ``ValueError`` raised by the broken DataFrame constructor (:401-402),
    which happens after every image product is complete.
"""
from __future__ import annotations

import builtins
import contextlib
import io
import os
import runpy
import sys
import types

import numpy as np

REFERENCE_SCRIPT = "/root/reference/cmf/robust_mf.py"

_ENVI_TO_NP = {"1": "u1", "2": "i2", "3": "i4", "4": "f4", "5": "f8", "12": "u2"}
_NP_CHAR_TO_ENVI = {"B": "1", "h": "2", "i": "3", "l": "3", "f": "4", "d": "5", "H": "12"}


def parse_envi_header(path):
    """Header text -> dict (lower-case keys, scalar -> str, ``{a, b}`` -> list of str)."""
    with open(path, "r") as fh:
        text = fh.read()
    lines = text.splitlines()
    if not lines or not lines[0].strip().upper().startswith("ENVI"):
        raise ValueError("%s is not an ENVI header" % path)
    meta = {}
    i = 1
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
            inner = val[val.index("{") + 1: val.rindex("}")]
            if key == "description":
                meta[key] = inner.strip()
            else:
                meta[key] = [tok.strip() for tok in inner.replace("\n", " ").split(",")]
        else:
            meta[key] = val
    return meta


def write_envi_header(path, meta):
    """dict -> header text.  Lists are written as ``{ a , b }``; strings verbatim."""
    first = ["samples", "lines", "bands", "header offset", "file type", "data type",
             "interleave", "byte order"]
    keys = [k for k in first if k in meta] + [k for k in meta if k not in first]
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


class _FakeImage(object):
    def __init__(self, hdr_path, data_path, meta):
        self.hdr_path, self.data_path, self.metadata = hdr_path, data_path, meta

    def _shape(self):
        m = self.metadata
        L, S, B = int(m["lines"]), int(m["samples"]), int(m["bands"])
        il = str(m["interleave"]).lower()
        return {"bil": (L, B, S), "bip": (L, S, B), "bsq": (B, L, S)}[il]

    def open_memmap(self, interleave="source", writable=False, writeable=None, **_):
        # the reference spells the keyword both ways (:207 vs :262)
        if writeable is not None:
            writable = writeable
        if interleave != "source":
            raise NotImplementedError("shim only serves interleave='source'")
        dt = np.dtype(_ENVI_TO_NP[str(self.metadata["data type"])])
        if str(self.metadata.get("byte order", "0")).strip() == "1":
            dt = dt.newbyteorder(">")
        off = int(self.metadata.get("header offset", 0))
        return np.memmap(self.data_path, dtype=dt, mode="r+" if writable else "r",
                         offset=off, shape=self._shape())


def _envi_open(hdr_path, image=None):
    meta = parse_envi_header(hdr_path)
    data = image if image is not None else hdr_path[:-4]
    return _FakeImage(hdr_path, data, meta)


def _envi_create_image(hdr_path, metadata, force=False, ext="", **_):
    data_path = hdr_path[:-4] + ext if hdr_path.endswith(".hdr") else hdr_path + ext
    if not force and (os.path.exists(hdr_path) or os.path.exists(data_path)):
        raise IOError("refusing to overwrite %s" % hdr_path)
    meta = dict(metadata)
    meta.setdefault("header offset", 0)
    meta.setdefault("byte order", 0)
    meta.setdefault("file type", "ENVI Standard")
    write_envi_header(hdr_path, meta)
    img = _FakeImage(hdr_path, data_path, {k: v for k, v in meta.items()})
    dt = np.dtype(_ENVI_TO_NP[str(meta["data type"])])
    nbytes = int(np.prod(img._shape())) * dt.itemsize
    with open(data_path, "wb") as fh:
        fh.truncate(nbytes)
    return img


def install_fake_spectral():
    """Put a minimal ``spectral.io.envi`` into sys.modules (idempotent)."""
    if "spectral.io.envi" in sys.modules and getattr(sys.modules["spectral.io.envi"], "_is_shim", False):
        return sys.modules["spectral.io.envi"]
    spectral = types.ModuleType("spectral")
    spectral_io = types.ModuleType("spectral.io")
    envi = types.ModuleType("spectral.io.envi")
    envi.open = _envi_open
    envi.create_image = _envi_create_image
    envi.dtype_to_envi = dict(_NP_CHAR_TO_ENVI)
    envi._is_shim = True
    spectral.io = spectral_io
    spectral_io.envi = envi
    sys.modules["spectral"] = spectral
    sys.modules["spectral.io"] = spectral_io
    sys.modules["spectral.io.envi"] = envi
    return envi


def import_reference_functions(script=REFERENCE_SCRIPT):
    """Return the reference's own module-level functions (cov, inv, det, eig, looshrinkage)."""
    install_fake_spectral()
    ns = runpy.run_path(script, run_name="robust_mf_reference")
    return types.SimpleNamespace(**{k: ns[k] for k in ("cov", "inv", "det", "eig", "looshrinkage")})


def run_reference_cli(argv, script=REFERENCE_SCRIPT, quiet=True):
    """Run the unmodified script's ``__main__`` body with ``sys.argv = ['robust_mf.py', *argv]``.

    Returns the captured stdout.  The known terminal ValueError (:401-402) is swallowed;
    anything else propagates.
    """
    install_fake_spectral()
    init = {
        "os": os, "sys": sys, "np": np,
        "map": lambda f, a: list(builtins.map(f, a)),
    }
    old_argv = sys.argv
    sys.argv = ["robust_mf.py"] + [str(a) for a in argv]
    buf = io.StringIO()
    try:
        ctx = contextlib.redirect_stdout(buf) if quiet else contextlib.nullcontext()
        with ctx:
            try:
                runpy.run_path(script, init_globals=init, run_name="__main__")
            except ValueError as exc:  # pandas shape mismatch at the very end
                if "Shape of passed values" not in str(exc) and "Length of values" not in str(exc) \
                        and "index" not in str(exc).lower():
                    raise
    finally:
        sys.argv = old_argv
    return buf.getvalue()


def write_bil_cube(path, cube_lbs, extra_meta=None, nodata=-9999):
    """Write a (lines, bands, samples) float32 array as an ENVI BIL file + header."""
    cube = np.ascontiguousarray(cube_lbs, dtype=np.float32)
    L, B, S = cube.shape
    cube.tofile(path)
    meta = {"samples": S, "lines": L, "bands": B, "header offset": 0,
            "file type": "ENVI Standard", "data type": 4, "interleave": "bil",
            "byte order": 0, "data ignore value": nodata}
    if extra_meta:
        meta.update(extra_meta)
    write_envi_header(path + ".hdr", meta)
    return path
