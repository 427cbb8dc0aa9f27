"""TEST INFRASTRUCTURE ONLY — loads the *live* reference for differential tests.

Imports `/root/reference/src/sparsification/{core,metrics,random,metric_backbone}.py`
unmodified, as a synthetic package, so the restatements in `oracle/` (and the
golden fixtures in `tests/golden/`) can be pinned against the reference itself.

* The reference's top-level `src/__init__.py` cannot be imported (it needs the
  absent `src/data`), so the four files are loaded as package `_gsr_reference`.
* `torch_geometric` is not installed in this image; a stub module exposing the
  repo's own `Data` duck-type is injected for `core.py:12` / `random.py:11`.
* `stable=True` rebinds the modules' `np` name to a proxy whose `argsort`
  forces `kind="stable"` (reference call sites `core.py:233,434,446`,
  `random.py:49`). The reference files are not edited. Default-kind argsort is
  an unstable SIMD sort on AVX-512 hosts, so the *reference's own* masks are
  platform dependent under ties; the stable variant is the parity contract.

`/root/reference` exists only in the build container: `available()` is False on
the GPU box and every caller must skip (never fail) in that case. Nothing on
the product path imports this module.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GSP_REFERENCE_ROOT", "/root/reference")
_PKG_DIR = os.path.join(REFERENCE_ROOT, "src", "sparsification")
_PKG_NAME = "_gsr_reference"


def available() -> bool:
    return os.path.isfile(os.path.join(_PKG_DIR, "core.py"))


def _ensure_pyg_stub() -> None:
    try:
        import torch_geometric.data  # noqa: F401  (real PyG wins when present)
        return
    except Exception:
        pass
    repo_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if repo_root not in sys.path:
        sys.path.insert(0, repo_root)
    from gsr_b200.data import Data

    tg = types.ModuleType("torch_geometric")
    tg_data = types.ModuleType("torch_geometric.data")
    tg_data.Data = Data
    tg.data = tg_data
    tg.__gsp_stub__ = True
    sys.modules["torch_geometric"] = tg
    sys.modules["torch_geometric.data"] = tg_data


class _StableNumpy:
    """Proxy for the numpy module whose argsort defaults to a stable sort."""

    def __init__(self, np_module):
        self.__dict__["_np"] = np_module

    def __getattr__(self, name):
        return getattr(self._np, name)

    def argsort(self, a, axis=-1, kind=None, order=None, **kw):
        return self._np.argsort(a, axis=axis, kind="stable", order=order, **kw)


def load(stable: bool = True):
    """Return the reference `src.sparsification` package (as `_gsr_reference[_stable]`)."""
    if not available():
        raise FileNotFoundError(f"reference not present at {_PKG_DIR}")
    name = _PKG_NAME + ("_stable" if stable else "")
    if name in sys.modules:
        return sys.modules[name]
    _ensure_pyg_stub()
    spec = importlib.util.spec_from_file_location(
        name, os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR]
    )
    pkg = importlib.util.module_from_spec(spec)
    sys.modules[name] = pkg
    spec.loader.exec_module(pkg)
    if stable:
        import numpy as np

        proxy = _StableNumpy(np)
        for sub in ("core", "random"):
            importlib.import_module(f"{name}.{sub}").np = proxy
    return pkg
