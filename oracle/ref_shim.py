"""TEST INFRASTRUCTURE ONLY -- compatibility shim that lets the *unmodified* reference
arithmetic modules under /root/reference import in this container (transformers 5.5.0, no
`diffusers`).  It is used by `oracle/make_golden.py` (golden-vector generation) and by the
`-m "not gpu"` cross-check tests when /root/reference exists; nothing in the product path,
`bench.py` or the `-m gpu` tests may import it (the reference tree is absent on the GPU box).

What it patches (all outside the reference tree, nothing is copied from it):
  * fake `diffusers` package providing the five names `vibevoice/schedule/dpm_solver.py:23-26`
    imports (ConfigMixin, register_to_config, SchedulerMixin, SchedulerOutput,
    KarrasDiffusionSchedulers, deprecate, randn_tensor);
  * `AutoModel.register` forced to `exist_ok=True` (transformers 5.5.0 ships its own
    `vibevoice_acoustic_tokenizer` model type, `modular_vibevoice_tokenizer.py:1188` collides);
  * alias for the removed `transformers.models.qwen2.tokenization_qwen2_fast`
    (`modular_vibevoice_text_tokenizer.py:7`).
"""
from __future__ import annotations

import functools
import inspect
import os
import sys
import types
from dataclasses import dataclass

REFERENCE_ROOT = os.environ.get("VV_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "vibevoice", "modular"))


def _install_fake_diffusers():
    if "diffusers" in sys.modules:
        return
    import torch

    class _Cfg(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError as e:  # pragma: no cover
                raise AttributeError(k) from e

    class ConfigMixin:
        config_name = None

        def register_to_config(self, **kw):
            if not hasattr(self, "_internal_dict"):
                self._internal_dict = _Cfg()
            self._internal_dict.update(kw)

        @property
        def config(self):
            return self._internal_dict

        @classmethod
        def from_config(cls, config, **kw):
            d = dict(config)
            d.update(kw)
            sig = inspect.signature(cls.__init__)
            return cls(**{k: v for k, v in d.items() if k in sig.parameters})

    def register_to_config(init):
        @functools.wraps(init)
        def wrapper(self, *args, **kwargs):
            sig = inspect.signature(init)
            bound = sig.bind(self, *args, **kwargs)
            bound.apply_defaults()
            cfg = {k: v for k, v in bound.arguments.items() if k != "self"}
            ConfigMixin.register_to_config(self, **cfg)
            init(self, *args, **kwargs)

        return wrapper

    class SchedulerMixin:
        pass

    @dataclass
    class SchedulerOutput:
        prev_sample: "torch.Tensor"

    import enum

    class KarrasDiffusionSchedulers(enum.Enum):
        DPMSolverMultistepScheduler = 1

    def deprecate(*a, **k):
        return None

    def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
        return torch.randn(shape, generator=generator, dtype=dtype).to(device)

    pkg = types.ModuleType("diffusers")
    cu = types.ModuleType("diffusers.configuration_utils")
    cu.ConfigMixin, cu.register_to_config = ConfigMixin, register_to_config
    ut = types.ModuleType("diffusers.utils")
    ut.deprecate = deprecate
    tu = types.ModuleType("diffusers.utils.torch_utils")
    tu.randn_tensor = randn_tensor
    sch = types.ModuleType("diffusers.schedulers")
    su = types.ModuleType("diffusers.schedulers.scheduling_utils")
    su.KarrasDiffusionSchedulers = KarrasDiffusionSchedulers
    su.SchedulerMixin, su.SchedulerOutput = SchedulerMixin, SchedulerOutput
    pkg.configuration_utils, pkg.utils, pkg.schedulers = cu, ut, sch
    ut.torch_utils, sch.scheduling_utils = tu, su
    for m in (pkg, cu, ut, tu, sch, su):
        sys.modules[m.__name__] = m


def _patch_transformers():
    from transformers.models.auto import auto_factory

    orig = auto_factory._LazyAutoMapping.register
    if not getattr(orig, "_vv_patched", False):

        def register(self, key, value, exist_ok=False):
            return orig(self, key, value, exist_ok=True)

        register._vv_patched = True
        auto_factory._LazyAutoMapping.register = register

    name = "transformers.models.qwen2.tokenization_qwen2_fast"
    if name not in sys.modules:
        try:
            from transformers.models.qwen2 import tokenization_qwen2 as tq

            mod = types.ModuleType(name)
            fast = getattr(tq, "Qwen2TokenizerFast", None) or getattr(tq, "Qwen2Tokenizer")
            mod.Qwen2TokenizerFast = fast
            sys.modules[name] = mod
        except Exception:  # pragma: no cover
            pass


_LOADED = {}


def load_reference():
    """Import the reference arithmetic modules; returns a namespace of the symbols the oracle
    is pinned against.  Raises RuntimeError when /root/reference is absent."""
    if _LOADED:
        return _LOADED["ns"]
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_fake_diffusers()
    _patch_transformers()
    # this repo ships a drop-in `vibevoice/` alias package; make sure the names below resolve to the REFERENCE tree
    for name in [n for n in sys.modules if n == "vibevoice" or n.startswith("vibevoice.")]:
        f = getattr(sys.modules[name], "__file__", "") or ""
        if not f.startswith(REFERENCE_ROOT):
            del sys.modules[name]
    if REFERENCE_ROOT in sys.path:
        sys.path.remove(REFERENCE_ROOT)
    sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    cfg = importlib.import_module("vibevoice.modular.configuration_vibevoice")
    head = importlib.import_module("vibevoice.modular.modular_vibevoice_diffusion_head")
    tok = importlib.import_module("vibevoice.modular.modular_vibevoice_tokenizer")
    dpm = importlib.import_module("vibevoice.schedule.dpm_solver")
    ns = types.SimpleNamespace(cfg=cfg, head=head, tok=tok, dpm=dpm)
    try:
        ns.modeling = importlib.import_module("vibevoice.modular.modeling_vibevoice")
    except Exception as e:  # pragma: no cover - only the connector lives here
        ns.modeling = None
        ns.modeling_error = repr(e)
    _LOADED["ns"] = ns
    return ns
