"""Drop-in import path for the reference's `vibevoice/modular/lora_loading.py` (`load_lora_assets`, :148-176)."""
from vibevoice_b200.lora import LoadReport as _LoadReport, load_lora_assets  # noqa: F401
