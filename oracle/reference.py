"""Access to the REAL reference build (oracle/_ref) -- TEST INFRASTRUCTURE ONLY.

In the build container /root/reference exists: ``load()`` makes ``import mettagrid`` resolve to the
reference's own Python package (via a scratch directory of symlinks, nothing is copied) together
with the pybind module compiled by oracle/Makefile.ref.  On the GPU box only the compiled module
travels; ``load_module_only()`` imports it as a top-level ``mettagrid_c`` and oracle/ref_driver.py
builds its GameConfig from our own config objects.
"""

from __future__ import annotations

import importlib
import os
import sys
import types
from pathlib import Path

_DIR = Path(__file__).resolve().parent
REF_ROOT = Path(os.environ.get("METTAGRID_REFERENCE", "/root/reference"))
REF_BUILD = _DIR / "_ref"


def so_path() -> Path | None:
    hits = sorted(REF_BUILD.glob("mettagrid_c*.so"))
    return hits[0] if hits else None


def available() -> bool:
    return so_path() is not None and (REF_ROOT / "python" / "src" / "mettagrid").is_dir()


def load():
    """Import the reference Python package + compiled module.  Returns the ``mettagrid`` module or
    None when the reference tree is not present (e.g. on the GPU box)."""
    if not available():
        return None
    if "mettagrid" in sys.modules and getattr(sys.modules["mettagrid"], "__graft_ref__", False):
        return sys.modules["mettagrid"]
    pkg_root = REF_BUILD / "pkg"
    pkg = pkg_root / "mettagrid"
    pkg.mkdir(parents=True, exist_ok=True)
    for entry in (REF_ROOT / "python" / "src" / "mettagrid").iterdir():
        link = pkg / entry.name
        if not link.exists():
            link.symlink_to(entry)
    so = so_path()
    link = pkg / so.name
    if not link.exists():
        link.symlink_to(so)
    # minimal stand-in for a third-party module the reference imports but the step path never uses
    if "importnb" not in sys.modules:
        try:
            importlib.import_module("importnb")
        except ImportError:
            stub = types.ModuleType("importnb")

            class Notebook:  # noqa: D401
                def __init__(self, *a, **k):
                    pass

                @staticmethod
                def load_module(*a, **k):
                    raise ImportError("importnb is not installed")

            stub.Notebook = Notebook
            sys.modules["importnb"] = stub
    if str(pkg_root) not in sys.path:
        sys.path.insert(0, str(pkg_root))
    mod = importlib.import_module("mettagrid")
    mod.__graft_ref__ = True
    return mod


def load_module_only():
    """Import just the compiled reference module (no reference Python needed)."""
    so = so_path()
    if so is None:
        return None
    if "mettagrid_c" in sys.modules:
        return sys.modules["mettagrid_c"]
    spec = importlib.util.spec_from_file_location("mettagrid_c", so)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["mettagrid_c"] = mod
    spec.loader.exec_module(mod)
    return mod
