"""Stub-import harness for the reference's own classes.  TEST INFRASTRUCTURE ONLY.

Works only where /root/reference exists (the build container), never on the GPU
box.  The reference scripts keep all work under `if __name__ == "__main__"`, so
importing them is side-effect free once three absent modules are stubbed
(SURVEY.md section 8c): `faiss`, `matplotlib(.pyplot)` and `models.lstm`
(the reference imports `models.lstm.Model` at LstmDistillFromDinoV2Train.py:5 but does
not ship it).  Used by oracle/make_golden.py and tests/test_oracle_vs_reference.py.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CSN_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "LstmDistillFromDinoV2Train.py"))


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def import_reference(module_name: str, model_cls=None):
    """Import a top-level reference script/module (e.g. 'LstmDistillFromDinoV2Train',
    'LstmDistillation', 'utils.utils', 'utils.EEGFilters') with the absent deps stubbed.
    `model_cls` is what `models.lstm.Model` resolves to inside the reference module."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.get(k) for k in
                  ("faiss", "matplotlib", "matplotlib.pyplot", "models", "models.lstm", "utils", "dino",
                   "cv2", "torchvision", "torchvision.transforms", "torchvision.datasets", "sklearn")}
    # our own repo has top-level `models/` (drop-in shim) -- hide it while importing the reference
    for k in list(sys.modules):
        if k == "utils" or k.startswith("utils.") or k == "dino" or k.startswith("dino."):
            del sys.modules[k]
    try:
        sys.path.insert(0, REFERENCE_ROOT)
        if "faiss" not in sys.modules or sys.modules["faiss"] is None:
            _stub("faiss")
        try:
            importlib.import_module("matplotlib.pyplot")
        except Exception:
            mpl = _stub("matplotlib")
            plt = _stub("matplotlib.pyplot")
            mpl.pyplot = plt
        for name in ("cv2",):
            try:
                importlib.import_module(name)
            except Exception:
                _stub(name)
        try:
            importlib.import_module("torchvision")
        except Exception:
            tv = _stub("torchvision")
            tv.transforms = _stub("torchvision.transforms")
            tv.datasets = _stub("torchvision.datasets")
        models_pkg = _stub("models")
        lstm_mod = _stub("models.lstm", Model=model_cls if model_cls is not None else object)
        models_pkg.lstm = lstm_mod
        return importlib.import_module(module_name)
    finally:
        sys.path[:] = saved_path
        for k in ("models", "models.lstm"):
            if saved_mods.get(k) is not None:
                sys.modules[k] = saved_mods[k]
            else:
                sys.modules.pop(k, None)
