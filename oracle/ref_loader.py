"""TEST / BENCH INFRASTRUCTURE ONLY -- loads the UNMODIFIED reference modules of the hot path.

The reference (granularai/fabric) is pure Python; its hot path is three files:
``models/bidate_model.py``, ``models/unet_parts.py`` and ``utils/metrics.py``.  They are loaded here by FILE PATH under
private module names (``_fabric_ref.models...``), so neither this repo's ``models/`` pickle shim (a regular package,
which would win over the reference's namespace package on ``sys.path``) nor any ``utils`` package can shadow them.

Search order: ``$FABRIC_REFERENCE`` / ``/root/reference`` (the build container), then ``oracle/_ref/`` -- a git-ignored,
build-time copy of exactly those three files written by ``oracle/build_ref.py`` (called from
``__graft_entry__.build()``), which travels to the GPU box with the snapshot like a built ``.so``.  Nothing in the
product path (``fabric_b200/``) imports this module; only ``tests/``, ``oracle/make_golden.py`` and the reference arm /
``cpu_baseline`` leg of ``bench.py`` do.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = ("models/bidate_model.py", "models/unet_parts.py", "utils/metrics.py")
_PKG = "_fabric_ref"
_cache = {}


def find_root():
    """Directory holding the three reference files, or None."""
    for root in (os.environ.get("FABRIC_REFERENCE"), "/root/reference", os.path.join(HERE, "_ref")):
        if root and all(os.path.exists(os.path.join(root, f)) for f in FILES):
            return root
    return None


def available() -> bool:
    return find_root() is not None


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """(BiDateNet class, metrics module, unet_parts module, root) of the unmodified reference."""
    root = find_root()
    if root is None:
        raise FileNotFoundError("reference sources not found (/root/reference or oracle/_ref; run oracle/build_ref.py)")
    if root in _cache:
        return _cache[root]
    dont = sys.dont_write_bytecode
    sys.dont_write_bytecode = True          # /root/reference is read-only
    try:
        for pkg in (_PKG, _PKG + ".models", _PKG + ".utils"):
            if pkg not in sys.modules:
                m = types.ModuleType(pkg)
                m.__path__ = []             # a package, so the reference's relative import resolves
                sys.modules[pkg] = m
        parts = _load(_PKG + ".models.unet_parts", os.path.join(root, "models", "unet_parts.py"))
        net = _load(_PKG + ".models.bidate_model", os.path.join(root, "models", "bidate_model.py"))
        metrics = _load(_PKG + ".utils.metrics", os.path.join(root, "utils", "metrics.py"))
    finally:
        sys.dont_write_bytecode = dont
    _cache[root] = (net.BiDateNet, metrics, parts, root)
    return _cache[root]
