"""Model registry under the reference's names: engines call `load_model(cfg["model"]["name"])(**kwargs)`
(engine/forgery_engine.py:141; registry of the reference: model/__init__.py:7-17)."""
from . import unidefense as _ud

MODEL = {cls_name: getattr(_ud, attr) for cls_name, attr in (
    ("UDEB4", "UniDefenseModelEb4"),      # EfficientNet-B4 backbone
    ("UDR18", "UniDefenseModelRes18"),    # ResNet-18 backbone
    ("UDR50", "UniDefenseModelRes50"),    # ResNet-50 backbone
)}
UniDefenseModelEb4, UniDefenseModelRes18, UniDefenseModelRes50 = (MODEL[k] for k in ("UDEB4", "UDR18", "UDR50"))


def load_model(name="UDE"):
    """Case-insensitive lookup; an unknown name fails the same way the reference does (AssertionError)."""
    key = str(name).upper()
    if key not in MODEL:
        raise AssertionError(f"Model '{name}' not found.")
    print(f"Using model: '{name}'")
    return MODEL[key]
