"""Deterministic, name-keyed tensors shared by make_golden.py (reference side, run in the
build container) and the parity tests (product side, run anywhere).  Because every weight is
regenerated from its state_dict name, the fixtures only need to store activations and
outputs, not multi-MB state_dicts.  CPU generator => identical values on every machine."""
import zlib

import torch


def _gen(name: str, salt: int = 0) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) + 7919 * salt) & 0x7FFFFFFF)
    return g


def tensor_for(name: str, shape, kind: str = "auto", salt: int = 0, dtype=torch.float32) -> torch.Tensor:
    """kind: 'conv' (N(0, 1/fan_in)), 'gamma' (U[.5,1.5]), 'beta' (N(0,.1)), 'var' (U[.5,1.5]),
    'mean' (N(0,.2)), 'unit' (U[-1,1]), 'normal' (N(0,1)), 'coef' (0.3), 'auto' (by name/shape)."""
    shape = tuple(int(s) for s in shape)
    g = _gen(name, salt)
    if kind == "auto":
        if name.endswith("fuse_coef") or name.endswith("sf_coef"):
            kind = "coef"
        elif name.endswith("running_mean"):
            kind = "mean"
        elif name.endswith("running_var"):
            kind = "var"
        elif name.endswith("num_batches_tracked"):
            return torch.zeros(shape, dtype=torch.int64)
        elif len(shape) >= 2:
            kind = "conv"
        elif name.endswith("bias"):
            kind = "beta"
        else:
            kind = "gamma"
    if kind == "coef":
        return torch.full(shape, 0.3, dtype=dtype)
    if kind == "conv":
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return (torch.randn(shape, generator=g, dtype=torch.float32) / max(fan_in, 1) ** 0.5).to(dtype)
    if kind in ("gamma", "var"):
        return (torch.rand(shape, generator=g, dtype=torch.float32) + 0.5).to(dtype)
    if kind == "beta":
        return (torch.randn(shape, generator=g, dtype=torch.float32) * 0.1).to(dtype)
    if kind == "mean":
        return (torch.randn(shape, generator=g, dtype=torch.float32) * 0.2).to(dtype)
    if kind == "unit":
        return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1).to(dtype)
    if kind == "normal":
        return torch.randn(shape, generator=g, dtype=torch.float32).to(dtype)
    raise ValueError(kind)


def fill_state_dict_(module: torch.nn.Module, prefix_filter=None, salt: int = 0):
    """Overwrite params/buffers of `module` (optionally only names passing prefix_filter)."""
    sd = module.state_dict()
    with torch.no_grad():
        for name, t in sd.items():
            if prefix_filter is not None and not prefix_filter(name):
                continue
            t.copy_(tensor_for(name, t.shape, salt=salt).to(t.dtype))
    return module


HOT_PREFIXES = ("dec_block", "freq_filter", "spat_filter", "bottleneck", "classifier", "fuse_coef")


def is_hot(name: str) -> bool:
    return name.startswith(HOT_PREFIXES)


def sample_indices(numel: int, k: int = 64, name: str = "") -> torch.Tensor:
    g = _gen("idx:" + name)
    if numel <= k:
        return torch.arange(numel)
    return torch.randint(0, numel, (k,), generator=g)
