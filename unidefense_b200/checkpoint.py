"""Checkpoint interchange with the reference engines.

The reference saves `{"step", "best_step", "best_auc", "best_acc", "model": model_without_ddp.state_dict()}` as
`best_model.bin` / `latest_model.bin` (engine/forgery_engine.py:215-223, ocim_engine.py:202-210,
uniattack_engine.py:258-268; the UniAttack engine stores `best_auc_frame` / `best_hter_frame` instead) and its test
path does `model.load_state_dict(ckpt["model"])` strictly (forgery_engine.py:200-209).  The drop-in models keep every
`state_dict` key, shape and the key ORDER of the reference classes (tests/test_model_api.py), so the files interchange;
this module is the checked way to do it:

* `load_reference_checkpoint(model, path)`  -- strict load of a reference file into a drop-in model; accepts a bare
  `state_dict`, a DDP-wrapped one (`module.` prefix) and a SyncBatchNorm-converted one; reports every missing /
  unexpected / mis-shaped key at once instead of torch's first-error message; returns the file's metadata.
* `save_reference_checkpoint(model, path, ...)` -- writes the reference's dict; tensors are stored contiguous in the
  default (NCHW) memory format, whatever layout the live parameters use (the 3x3 filter weight of the tcgen05
  projection lives channels-last in memory), so a stock-torch reader sees exactly the reference's file.
* `verify(model, path)` -- load -> save -> reload round trip, bit-exact.

No kernels are involved: this runs on any device."""
import io
import os
from collections import OrderedDict

import torch

META_KEYS = ("step", "best_step", "best_auc", "best_acc", "best_auc_frame", "best_hter_frame", "best_auc_video",
             "best_hter_video", "best_hter")


def _unwrap(model):
    return model.module if hasattr(model, "module") and isinstance(model.module, torch.nn.Module) else model


def _read(path_or_file):
    try:
        ckpt = torch.load(path_or_file, map_location="cpu", weights_only=True)
    except Exception:        # noqa: BLE001 -- older files pickle numpy scalars in the metadata
        if hasattr(path_or_file, "seek"):
            path_or_file.seek(0)
        ckpt = torch.load(path_or_file, map_location="cpu", weights_only=False)
    if isinstance(ckpt, dict) and "model" in ckpt and isinstance(ckpt["model"], dict):
        return ckpt["model"], {k: v for k, v in ckpt.items() if k != "model"}
    if isinstance(ckpt, dict) and ckpt and all(torch.is_tensor(v) for v in ckpt.values()):
        return ckpt, {}
    raise ValueError("not a UniDefense checkpoint: expected {'model': state_dict, ...} or a bare state_dict")


def _strip_prefix(sd):
    if sd and all(k.startswith("module.") for k in sd):
        return OrderedDict((k[len("module."):], v) for k, v in sd.items())
    return sd


def compare(model, state_dict):
    """-> dict(missing=[...], unexpected=[...], shape=[(key, ours, theirs)], order_equal=bool) for a drop-in model."""
    ours = _unwrap(model).state_dict()
    theirs = _strip_prefix(state_dict)
    missing = [k for k in ours if k not in theirs]
    unexpected = [k for k in theirs if k not in ours]
    shape = [(k, tuple(ours[k].shape), tuple(theirs[k].shape)) for k in ours
             if k in theirs and tuple(ours[k].shape) != tuple(theirs[k].shape)]
    return dict(missing=missing, unexpected=unexpected, shape=shape, order_equal=list(ours) == list(theirs))


def load_reference_checkpoint(model, path_or_file, strict=True):
    """Loads a reference `best_model.bin` (or a bare / DDP-prefixed state_dict) into `model`; returns its metadata."""
    sd, meta = _read(path_or_file)
    sd = _strip_prefix(sd)
    rep = compare(model, sd)
    if strict and (rep["missing"] or rep["unexpected"] or rep["shape"]):
        lines = [f"checkpoint does not match {type(_unwrap(model)).__name__}:"]
        lines += [f"  missing in file: {k}" for k in rep["missing"][:20]]
        lines += [f"  unexpected in file: {k}" for k in rep["unexpected"][:20]]
        lines += [f"  shape of {k}: model {a} vs file {b}" for k, a, b in rep["shape"][:20]]
        raise RuntimeError("\n".join(lines))
    skip = {k for k, _, _ in rep["shape"]}
    _unwrap(model).load_state_dict({k: v for k, v in sd.items() if k not in skip}, strict=strict)
    return meta


def reference_state_dict(model):
    """state_dict with every tensor on the CPU, contiguous in the default memory format, keys in the reference order."""
    out = OrderedDict()
    for k, v in _unwrap(model).state_dict().items():
        out[k] = v.detach().to("cpu").contiguous(memory_format=torch.contiguous_format).clone()
    return out


def save_reference_checkpoint(model, path_or_file, step=0, best_step=0, best_auc=0.0, best_acc=0.0, **extra):
    """Writes the dict the reference engines write (forgery_engine.py:215-223); `extra` adds e.g. best_auc_frame."""
    ckpt = {"step": step, "best_step": best_step, "best_auc": best_auc, "best_acc": best_acc}
    ckpt.update(extra)
    ckpt["model"] = reference_state_dict(model)
    torch.save(ckpt, path_or_file)


def verify(model, path_or_file):
    """load -> save -> reload; raises unless every tensor survives bit for bit.  Returns (metadata, n_tensors)."""
    meta = load_reference_checkpoint(model, path_or_file, strict=True)
    if hasattr(path_or_file, "seek"):
        path_or_file.seek(0)
    src, _ = _read(path_or_file)
    src = _strip_prefix(src)
    buf = io.BytesIO()
    save_reference_checkpoint(model, buf, **{k: v for k, v in meta.items() if k in META_KEYS})
    buf.seek(0)
    back, meta2 = _read(buf)
    if list(back) != list(src):
        raise RuntimeError("key order changed in the round trip")
    for k in src:
        if not torch.equal(back[k], src[k].to(back[k].dtype)) or not back[k].is_contiguous():
            raise RuntimeError(f"tensor {k} changed in the round trip")
    for k in meta:
        if k in META_KEYS and meta2.get(k) != meta[k]:
            raise RuntimeError(f"metadata {k} changed in the round trip")
    return meta, len(src)


def main(argv=None):
    import argparse
    import json
    ap = argparse.ArgumentParser(description="check / convert a UniDefense checkpoint against the drop-in models")
    ap.add_argument("command", choices=["verify", "inspect", "resave"])
    ap.add_argument("checkpoint")
    ap.add_argument("--model", default="UDEB4", help="UDEB4 / UDR18 / UDR50")
    ap.add_argument("--kwargs", default="{}", help='JSON of the YAML `model:` kwargs, e.g. \'{"num_classes": 2}\'')
    ap.add_argument("--out", default=None, help="resave: output file")
    a = ap.parse_args(argv)
    from unidefense_b200.model import load_model
    kw = json.loads(a.kwargs)
    if a.model == "UDEB4":
        kw.setdefault("extractor", "efficientnet-b4")
    model = load_model(a.model)(**kw)
    if a.command == "inspect":
        sd, meta = _read(a.checkpoint)
        rep = compare(model, sd)
        print(json.dumps({"metadata": {k: (v if isinstance(v, (int, float, str)) else str(v)) for k, v in meta.items()},
                          "tensors": len(sd), **{k: (v[:20] if isinstance(v, list) else v) for k, v in rep.items()}}, indent=1))
        return 0 if not (rep["missing"] or rep["unexpected"] or rep["shape"]) else 1
    meta, n = verify(model, a.checkpoint)
    print(f"{os.path.basename(a.checkpoint)}: {n} tensors load strictly into {a.model} and survive save -> reload bit for bit; "
          f"metadata {meta}")
    if a.command == "resave":
        if not a.out:
            ap.error("resave needs --out")
        save_reference_checkpoint(model, a.out, **{k: v for k, v in meta.items() if k in META_KEYS})
        print(f"wrote {a.out}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
