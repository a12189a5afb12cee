"""Checkpoint interchange with the reference engines: unidefense_b200.checkpoint.

CPU tests.  The cross-load cases build the LIVE reference classes (tests/golden/ref_loader.py) and are skipped where
/root/reference does not exist (the GPU box); the self round trip runs everywhere."""
import io
import os
import sys

import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLDEN)
import ref_loader  # noqa: E402

CTOR = {"r18": ("UDR18", dict(num_classes=2, drop_rate=0.5)),
        "r50": ("UDR50", dict(extractor="resnet50", num_classes=2, drop_rate=0.5)),
        "eb4": ("UDEB4", dict(extractor="efficientnet-b4", num_classes=2, drop_rate=0.2))}


def _ours(arch):
    from unidefense_b200.model import load_model
    name, kw = CTOR[arch]
    return load_model(name)(**kw)


def _randomise(model, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for v in model.state_dict().values():
            if v.dtype.is_floating_point:
                v.copy_(torch.randn(v.shape, generator=g))
            else:
                v.fill_(seed)


@pytest.mark.parametrize("arch", ["r18", "eb4"])
def test_self_round_trip_and_layout(arch):
    from unidefense_b200 import checkpoint as CK
    a = _ours(arch)
    _randomise(a, 3)
    buf = io.BytesIO()
    CK.save_reference_checkpoint(a, buf, step=7, best_step=5, best_auc=0.91, best_acc=0.88)
    buf.seek(0)
    raw = torch.load(buf, map_location="cpu", weights_only=True)
    assert list(raw) == ["step", "best_step", "best_auc", "best_acc", "model"]          # forgery_engine.py:217-223
    assert all(v.is_contiguous() for v in raw["model"].values())                         # also the channels-last 3x3 weight
    assert list(raw["model"]) == list(a.state_dict())
    b = _ours(arch)
    buf.seek(0)
    meta = CK.load_reference_checkpoint(b, buf)
    assert meta == dict(step=7, best_step=5, best_auc=0.91, best_acc=0.88)
    for (k, u), (k2, v) in zip(a.state_dict().items(), b.state_dict().items()):
        assert k == k2 and torch.equal(u, v), k
    buf.seek(0)
    assert CK.verify(b, buf)[1] == len(a.state_dict())


def test_ddp_prefix_bare_state_dict_and_errors():
    from unidefense_b200 import checkpoint as CK
    a, b = _ours("r18"), _ours("r18")
    _randomise(a, 1)
    buf = io.BytesIO()
    torch.save({"module." + k: v for k, v in a.state_dict().items()}, buf)              # bare + DDP-prefixed
    buf.seek(0)
    assert CK.load_reference_checkpoint(b, buf) == {}
    assert all(torch.equal(u, v) for u, v in zip(a.state_dict().values(), b.state_dict().values()))
    sd = dict(a.state_dict())
    sd.pop("fuse_coef")
    sd["classifier.fc.weight"] = torch.zeros(3, 512)
    sd["not_a_key"] = torch.zeros(1)
    buf = io.BytesIO()
    torch.save({"model": sd, "best_step": 1}, buf)
    buf.seek(0)
    with pytest.raises(RuntimeError) as e:
        CK.load_reference_checkpoint(b, buf)
    msg = str(e.value)
    assert "missing in file: fuse_coef" in msg and "unexpected in file: not_a_key" in msg and "classifier.fc.weight" in msg
    with pytest.raises(ValueError):
        buf = io.BytesIO()
        torch.save({"weights": 3}, buf)
        buf.seek(0)
        CK.load_reference_checkpoint(b, buf)


@pytest.mark.skipif(not ref_loader.available(), reason="needs the live reference at /root/reference")
@pytest.mark.parametrize("arch", ["r18", "r50", "eb4"])
def test_cross_load_with_the_live_reference_classes(arch, tmp_path):
    """A file written the way the reference engine writes it loads strictly into the drop-in, and a file written by
    the drop-in loads strictly into the reference class (forgery_engine.py:200-209), tensors bit-identical."""
    from unidefense_b200 import checkpoint as CK
    ref = ref_loader.load()
    name, kw = CTOR[arch]
    kw = dict(kw)
    torch.manual_seed(11)
    rmodel = ref.model.load_model(name)(**kw)
    _randomise(rmodel, 5)
    path = os.path.join(tmp_path, "best_model.bin")
    torch.save({"step": 12, "best_step": 10, "best_auc": 0.97, "best_acc": 0.93, "model": rmodel.state_dict()}, path)
    ours = _ours(arch)
    meta = CK.load_reference_checkpoint(ours, path)
    assert meta["best_step"] == 10
    for (k, u), (k2, v) in zip(rmodel.state_dict().items(), ours.state_dict().items()):
        assert k == k2 and u.shape == v.shape and torch.equal(u, v), k
    assert CK.verify(ours, path)[1] == len(rmodel.state_dict())
    # and back: the drop-in's file through the reference's own loading code
    _randomise(ours, 9)
    back = os.path.join(tmp_path, "latest_model.bin")
    CK.save_reference_checkpoint(ours, back, step=13, best_step=10, best_auc=0.97, best_acc=0.93)
    ckpt = torch.load(back, map_location="cpu")
    rmodel.load_state_dict(ckpt["model"])                                               # strict, as the engine does
    for u, v in zip(rmodel.state_dict().values(), ours.state_dict().values()):
        assert torch.equal(u, v)
    assert round(ckpt.get("best_auc", -1), 4) == 0.97                                   # the engine's print statement
