"""Fine-tuned adapter assets -> merged base weights (reference `vibevoice/modular/lora_loading.py:148-176`).

The reference wraps `model.model.language_model` and `model.model.prediction_head` in `peft.PeftModel` (:102-106, :129-137) and
keeps the low-rank pair next to every targeted `nn.Linear`, paying an extra pair of skinny GEMVs per projection per token.  The
generation path here packs weights once into kernel layouts, so adapters are folded into the base matrices **before** packing:

    W' = W + s * (B @ A),   s = lora_alpha / r   (PEFT LoRA; `use_rslora` -> lora_alpha / sqrt(r); `fan_in_fan_out` -> (B @ A)^T)

which is what `peft` (a `pyproject.toml` dependency of the reference, not vendored and not installed in this image) computes in
`LoraLayer.get_delta_weight` / `merge_and_unload`.  The merge is done in fp32 and rounded once to the storage dtype.

Directory layout read (same as the reference's loader):

    <ckpt>/lora/                         (or <ckpt> itself, :46-55)
        adapter_config.json + adapter_model.{safetensors,bin}          LoRA on the Qwen2 decoder (:116-127)
        diffusion_head/adapter_config.json + adapter_model.*           LoRA on the prediction head (:72-95)
        diffusion_head/diffusion_head_full.bin | diffusion_head_full.bin   full head state dict (:97-112)
        acoustic_connector/pytorch_model.bin, semantic_connector/pytorch_model.bin   full connector state dicts (:164-170)
"""
from __future__ import annotations

import json
import math
import re
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, Iterable, Iterator, Optional, Tuple

import torch


@dataclass
class LoadReport:
    """Mirror of the reference's `_LoadReport` (:17-27)."""
    language_model: bool = False
    diffusion_head_lora: bool = False
    diffusion_head_full: bool = False
    acoustic_connector: bool = False
    semantic_connector: bool = False
    adapter_root: Optional[Path] = None


def resolve_adapter_root(checkpoint_path) -> Path:
    """:46-55."""
    p = Path(checkpoint_path)
    if p.is_file():
        p = p.parent
    return p / "lora" if (p / "lora").exists() else p


def _read_tensors(stem: Path) -> Optional[Dict[str, torch.Tensor]]:
    st, bn = stem.with_suffix(".safetensors"), stem.with_suffix(".bin")
    if st.exists():
        from safetensors.torch import load_file
        return load_file(str(st), device="cpu")
    if bn.exists():
        return torch.load(str(bn), map_location="cpu", weights_only=True)
    return None


_LORA_KEY = re.compile(r"^(?:base_model\.model\.)?(?P<mod>.+?)\.lora_(?P<ab>[AB])(?:\.(?P<adapter>[^.]+))?\.weight$")


def lora_deltas(adapter_dir: Path, strip_prefix: str = "") -> Dict[str, torch.Tensor]:
    """{module path (relative to the wrapped module) : fp32 delta of its `.weight`} for one PEFT LoRA adapter directory.

    `strip_prefix` removes the attribute the reference's forward shim inserts (`base.` for the diffusion head, :30-43)."""
    cfg_path = adapter_dir / "adapter_config.json"
    tensors = _read_tensors(adapter_dir / "adapter_model")
    if not cfg_path.exists() or tensors is None:
        return {}
    cfg = json.loads(cfg_path.read_text())
    if cfg.get("peft_type", "LORA") != "LORA":
        raise NotImplementedError("adapter type %r (only LoRA adapters are produced by the reference's fine-tuning)" % cfg.get("peft_type"))
    if cfg.get("use_dora", False):
        raise NotImplementedError("DoRA adapters carry a magnitude vector and do not fold into a plain matrix")
    r_default, alpha_default = int(cfg["r"]), float(cfg.get("lora_alpha", cfg["r"]))
    rank_pat, alpha_pat = cfg.get("rank_pattern") or {}, cfg.get("alpha_pattern") or {}
    fan_in_fan_out = bool(cfg.get("fan_in_fan_out", False))
    pairs: Dict[str, Dict[str, torch.Tensor]] = {}
    for k, t in tensors.items():
        m = _LORA_KEY.match(k)
        if not m:
            if "lora_embedding" in k or "lora_magnitude" in k:
                raise NotImplementedError("adapter tensor %s: embedding / magnitude adapters are off this path" % k)
            continue
        mod = m.group("mod")
        if strip_prefix and mod.startswith(strip_prefix):
            mod = mod[len(strip_prefix):]
        pairs.setdefault(mod, {})[m.group("ab")] = t.float()
    out = {}
    for mod, ab in pairs.items():
        if "A" not in ab or "B" not in ab:
            raise ValueError("adapter for %s lacks lora_%s" % (mod, "B" if "A" in ab else "A"))
        A, B = ab["A"], ab["B"]                                   # A [r, in], B [out, r]
        r = A.shape[0]
        if B.shape[1] != r:
            raise ValueError("adapter for %s: rank mismatch %s vs %s" % (mod, tuple(A.shape), tuple(B.shape)))
        alpha = alpha_default
        for pat, v in alpha_pat.items():
            if re.search(r"(^|\.)%s$" % re.escape(pat), mod):
                alpha = float(v)
        r_cfg = r_default
        for pat, v in rank_pat.items():
            if re.search(r"(^|\.)%s$" % re.escape(pat), mod):
                r_cfg = int(v)
        s = alpha / math.sqrt(r_cfg) if cfg.get("use_rslora", False) else alpha / r_cfg
        d = (B @ A) * s
        out[mod] = d.T.contiguous() if fan_in_fan_out else d
    return out


def collect_overrides(checkpoint_dir) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor], LoadReport]:
    """(deltas, replacements, report): `deltas[key]` is added to the base tensor `key`, `replacements[key]` replaces it.
    Keys are full reference state-dict names (`model.language_model.…`, `model.prediction_head.…`, `model.*_connector.…`)."""
    root = resolve_adapter_root(checkpoint_dir)
    if not root.exists():
        raise FileNotFoundError("Adapter directory not found: %s" % root)                              # :160-161
    rep = LoadReport(adapter_root=root)
    deltas: Dict[str, torch.Tensor] = {}
    repl: Dict[str, torch.Tensor] = {}
    lm = lora_deltas(root)                                                                             # :116-137
    for mod, d in lm.items():
        deltas["model.language_model.%s.weight" % mod] = d
    rep.language_model = bool(lm)
    head_dir = root / "diffusion_head"
    hd = lora_deltas(head_dir, strip_prefix="base.") if head_dir.exists() else {}                      # :72-95
    for mod, d in hd.items():
        deltas["model.prediction_head.%s.weight" % mod] = d
    rep.diffusion_head_lora = bool(hd)
    if not hd:                                                                                         # :97-112 (fallback only)
        full = head_dir / "diffusion_head_full.bin"
        if not full.exists():
            full = root / "diffusion_head_full.bin"
        if full.exists():
            for k, t in torch.load(str(full), map_location="cpu", weights_only=True).items():
                repl["model.prediction_head.%s" % k] = t
            rep.diffusion_head_full = True
    for name in ("acoustic_connector", "semantic_connector"):                                          # :164-170
        p = root / name / "pytorch_model.bin"
        if p.exists():
            for k, t in torch.load(str(p), map_location="cpu", weights_only=True).items():
                repl["model.%s.%s" % (name, k)] = t
            setattr(rep, name, True)
    return deltas, repl, rep


def merged_state_dict(items: Iterable[Tuple[str, torch.Tensor]], deltas: Dict[str, torch.Tensor],
                      repl: Dict[str, torch.Tensor]) -> Iterator[Tuple[str, torch.Tensor]]:
    """Stream the base (name, tensor) pairs with the adapter assets applied; raises if an asset names a tensor the base lacks
    (the reference's `load_state_dict(strict=False)` only warns about those, :63-66 -- a silently ignored adapter is a wrong model)."""
    used = set()
    for name, t in items:
        if name in repl:
            used.add(name)
            if tuple(repl[name].shape) != tuple(t.shape):
                raise ValueError("replacement for %s has shape %s, base %s" % (name, tuple(repl[name].shape), tuple(t.shape)))
            t = repl[name].to(t.dtype)
        if name in deltas:
            used.add(name)
            d = deltas[name]
            if tuple(d.shape) != tuple(t.shape):
                raise ValueError("LoRA delta for %s has shape %s, base %s" % (name, tuple(d.shape), tuple(t.shape)))
            t = (t.float() + d).to(t.dtype)
        yield name, t
    missing = (set(deltas) | set(repl)) - used
    if missing:
        raise KeyError("adapter assets target tensors that the checkpoint does not have: %s" % sorted(missing)[:8])


def load_lora_assets(model, checkpoint_dir: str, device=None) -> LoadReport:
    """Drop-in for the reference's `load_lora_assets(model, checkpoint_dir, device)` (:148-176): folds the assets into the base
    weights and re-packs them.  The model must have been built by `from_pretrained` / `load_state_dict` from a re-readable
    source (it is streamed again; packed kernel layouts cannot be un-packed)."""
    deltas, repl, rep = collect_overrides(checkpoint_dir)
    if not (deltas or repl):
        import warnings
        warnings.warn("No adapter assets were loaded. Ensure the checkpoint directory is correct and contains LoRA weights.")   # :171-174
        return rep
    model._reload_with(lambda items: merged_state_dict(items, deltas, repl))
    return rep
