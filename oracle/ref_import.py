"""TEST INFRASTRUCTURE ONLY -- loader for the *real* reference implementation.

Imports ``/root/reference/benchmark/wifi_csi`` (model/that.py, utils.py, load_data.py)
with stub modules for the three packages the reference imports but this image lacks
(``ptflops``, ``matplotlib``, ``seaborn``) and with wandb disabled.  It exists so that

  * ``oracle/make_golden.py`` can generate the committed fixtures under ``tests/golden/``
  * the CPU tests can pin ``oracle/that_oracle.py`` (the restatement that travels to
    the GPU box) against the unmodified reference when ``/root/reference`` is present.

Nothing in the product package imports this file.  ``/root/reference`` does not exist on
the GPU box, so everything here is gated on :func:`reference_available`.
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("CSI_REFERENCE_ROOT", "/root/reference")
REF_WIFI = os.path.join(REF_ROOT, "benchmark", "wifi_csi")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_WIFI, "model", "that.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


_CACHE = {}


def load_reference():
    """Returns a namespace with the reference's ``that`` / ``utils`` / ``load_data`` / ``preset`` modules."""
    if "ns" in _CACHE:
        return _CACHE["ns"]
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    os.environ.setdefault("WANDB_MODE", "disabled")
    os.environ.setdefault("WANDB_SILENT", "true")
    _stub("ptflops", get_model_complexity_info=lambda *a, **k: (0, 0))
    mpl = _stub("matplotlib")
    plt = _stub("matplotlib.pyplot")
    mpl.pyplot = plt
    _stub("seaborn")
    if REF_WIFI not in sys.path:
        sys.path.insert(0, REF_WIFI)
    import torch
    prec = torch.get_float32_matmul_precision()
    mods = {}
    for name, rel in (("preset", "preset.py"), ("utils", "utils.py"), ("load_data", "load_data.py"),
                      ("train", "train.py")):
        # the reference modules import each other by bare name (``from preset import preset``)
        if name in sys.modules and getattr(sys.modules[name], "__file__", "").startswith(REF_WIFI):
            mods[name] = sys.modules[name]
            continue
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF_WIFI, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    spec = importlib.util.spec_from_file_location("ref_model_that", os.path.join(REF_WIFI, "model", "that.py"))
    that = importlib.util.module_from_spec(spec)
    sys.modules["ref_model_that"] = that
    spec.loader.exec_module(that)
    # train.py:24 sets "high" at import time; the oracle runs at full fp32 precision.
    torch.set_float32_matmul_precision(prec)
    ns = types.SimpleNamespace(that=that, **mods)
    _CACHE["ns"] = ns
    return ns
