"""Loads the REFERENCE's own modules (hippie/model.py, hippie/backbones.py, hippie/dataloading.py), unmodified.

TEST / BASELINE INFRASTRUCTURE ONLY: used by oracle/make_golden.py (pinning the oracle), by tests that compare the
oracle with the reference, and by the reference arm / `cpu_baseline` leg of bench.py.  Never imported by the product
(hippie_b200/, hippie/, scripts/).

Where the reference comes from
  * `baseline/_ref/`  -- the offline install of the unmodified reference package
        python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
               --target baseline/_ref <copy of /root/reference>
    made by `__graft_entry__.build()` in the build container (git-ignored, travels to the GPU box), or
  * `/root/reference` (build container only).

The reference has no `hippie/__init__.py` (a namespace package) while this repository ships a regular `hippie/` alias
package that would shadow it on any sys.path order.  The modules are therefore imported with `sys.modules["hippie"]`
temporarily replaced by a stand-in package whose `__path__` is the reference directory, so that the reference's own
`from hippie.backbones import ...` (hippie/model.py:6) resolves to ITS backbones; afterwards the entries are removed
again and the alias package is restored.  `pytorch_lightning` is not installed (no network): the 10-line stand-in of
oracle/_plstub supplies `LightningModule`.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CANDIDATES = (os.path.join(ROOT, "baseline", "_ref"), "/root/reference")

_cache = {}


def reference_root():
    """First directory that holds the reference's hippie/model.py, or None."""
    for c in CANDIDATES:
        if os.path.isfile(os.path.join(c, "hippie", "model.py")):
            return c
    return None


def load_reference(root: str | None = None):
    """-> namespace with .model, .backbones, .dataloading (the reference's modules) and .root."""
    root = root or reference_root()
    if root is None:
        raise FileNotFoundError("reference not found: neither baseline/_ref (run __graft_entry__.build() in the build "
                                "container) nor /root/reference holds hippie/model.py")
    if root in _cache:
        return _cache[root]
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "hippie" or k.startswith("hippie.")}
    had_pl = "pytorch_lightning" in sys.modules
    stub = os.path.join(HERE, "_plstub")
    pkg = types.ModuleType("hippie")
    pkg.__path__ = [os.path.join(root, "hippie")]
    sys.modules["hippie"] = pkg
    sys.path.insert(0, stub)
    try:
        mods = {n: importlib.import_module("hippie." + n) for n in ("backbones", "model", "dataloading")}
    finally:
        sys.path.remove(stub)
        for k in [k for k in sys.modules if k == "hippie" or k.startswith("hippie.")]:
            del sys.modules[k]
        sys.modules.update(saved)
        if not had_pl:  # the stand-in must not leak into code that probes for real Lightning
            for k in [k for k in sys.modules if k == "pytorch_lightning" or k.startswith("pytorch_lightning.")]:
                del sys.modules[k]
    for n, m in mods.items():
        assert os.path.abspath(m.__file__).startswith(os.path.abspath(root)), (n, m.__file__)
    ns = types.SimpleNamespace(root=root, **mods)
    _cache[root] = ns
    return ns
