"""Import the UniDefense reference (read-only at /root/reference) in THIS container.

Test infrastructure only.  The reference needs `timm` and `matplotlib`, which are not
installed; only three timm callables are ever *invoked* with the shipped configs
(SURVEY.md App. C), so tiny `sys.modules` stubs are enough.  Nothing here is used at
run time on the GPU box (where /root/reference does not exist): this module only feeds
`make_golden.py`, which writes the small fixtures under tests/golden/.
"""
import importlib.util
import os
import sys
import types

import torch.nn as nn

REF_ROOT = os.environ.get("UD_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "model"))


def _install_stubs():
    if "timm" in sys.modules and getattr(sys.modules["timm"], "_ud_stub", False):
        return
    timm = types.ModuleType("timm")
    timm._ud_stub = True
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    helpers = types.ModuleType("timm.models.helpers")

    class _Unused(nn.Module):  # drop_block / drop_path / aa_layer are off in every template
        def __init__(self, *a, **k):
            raise RuntimeError("timm stub: this layer is never built by the shipped configs")

    for name in ("DropBlock2d", "DropPath", "AvgPool2dSame", "BlurPool2d", "GroupNorm"):
        setattr(layers, name, _Unused)
    layers.create_attn = lambda attn_layer, planes: None
    layers.get_attn = lambda attn_layer: None

    def create_classifier(num_features, num_classes, pool_type="avg"):
        return nn.AdaptiveAvgPool2d(1), nn.Linear(num_features, max(num_classes, 1))

    layers.create_classifier = create_classifier

    def build_model_with_cfg(cls, variant, pretrained, **kw):
        kw.pop("pretrained_cfg", None)
        return cls(**kw)

    helpers.build_model_with_cfg = build_model_with_cfg
    helpers.checkpoint_seq = lambda fns, x, **k: x
    timm.models = models
    models.layers = layers
    models.helpers = helpers
    sys.modules.update({"timm": timm, "timm.models": models,
                        "timm.models.layers": layers, "timm.models.helpers": helpers})
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt})


_cached = {}


def load():
    """Returns a namespace with the reference's `model`, `loss`, `utils.operation` modules."""
    if "ns" in _cached:
        return _cached["ns"]
    if not available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    _install_stubs()
    # The reference uses top-level package names (model, loss, utils); make sure ours
    # cannot shadow them and that they resolve to /root/reference.
    for k in [k for k in sys.modules if k.split(".")[0] in ("model", "loss", "utils", "engine")]:
        del sys.modules[k]
    sys.path.insert(0, REF_ROOT)
    try:
        import model as ref_model                      # noqa
        import model.efficientnet.model as eff_model   # noqa
        eff_model.load_pretrained_weights = lambda *a, **k: None   # random init, no download
        import model.unidefense as ref_unidefense      # noqa
        import model.modules as ref_modules            # noqa
        import loss as ref_loss                        # noqa
        import utils.operation as ref_operation        # noqa
    finally:
        sys.path.remove(REF_ROOT)
    ns = types.SimpleNamespace(model=ref_model, unidefense=ref_unidefense, modules=ref_modules,
                               loss=ref_loss, operation=ref_operation,
                               efficientnet=sys.modules["model.efficientnet"],
                               resnet=sys.modules["model.resnet"])
    _cached["ns"] = ns
    return ns


def load_abstract_engine():
    """AbstractEngine loaded by file path (engine/__init__ pulls lmdb/albumentations)."""
    load()
    sys.path.insert(0, REF_ROOT)
    try:
        spec = importlib.util.spec_from_file_location(
            "ref_abstract_engine", os.path.join(REF_ROOT, "engine", "abstract_engine.py"))
        mod = importlib.util.module_from_spec(spec)
        wandb = sys.modules.get("wandb")
        if wandb is None:
            sys.modules["wandb"] = types.ModuleType("wandb")
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(REF_ROOT)
    return mod.AbstractEngine
