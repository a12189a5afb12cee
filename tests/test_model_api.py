"""Drop-in boundary: model registry, constructor kwargs, module tree constraints and
state_dict names/shapes equal to the reference's (recorded by make_golden.py from the reference classes)."""
import os

import pytest
import torch
import torch.nn as nn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CTOR = {"eb4": ("UDEB4", dict(extractor="efficientnet-b4", num_classes=2, drop_rate=0.2)),
        "r18": ("UDR18", dict(num_classes=2, drop_rate=0.5)),
        "r50": ("UDR50", dict(extractor="resnet50", num_classes=2, drop_rate=0.5))}


def _build(arch):
    from unidefense_b200.model import load_model
    name, kw = CTOR[arch]
    return load_model(name)(**kw)


@pytest.mark.parametrize("arch", ["eb4", "r18", "r50"])
def test_state_dict_matches_reference(arch):
    ref_shapes = torch.load(os.path.join(GOLDEN, f"full_{arch}.pt"), weights_only=False)["state_dict_shapes"]
    model = _build(arch)
    own = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert set(own) == set(ref_shapes), (sorted(set(own) - set(ref_shapes))[:5], sorted(set(ref_shapes) - set(own))[:5])
    assert own == ref_shapes
    assert list(own) == list(ref_shapes), "state_dict ordering differs (optimizer param order would too)"
    # frozen bottleneck bias, scalar fuse_coef at 0, sf_coef at -10 (Appendix D)
    assert not model.bottleneck.bias.requires_grad
    assert float(model.fuse_coef) == 0.0
    assert all(float(p) == -10.0 for n, p in model.named_parameters() if n.endswith("sf_coef"))
    assert model.path == "model/unidefense.py"


@pytest.mark.parametrize("arch", ["r18", "eb4"])
def test_module_tree_survives_syncbn_and_param_groups(arch):
    model = _build(arch)
    n_bn = sum(isinstance(m, nn.modules.batchnorm._BatchNorm) for m in model.modules())
    sync = nn.SyncBatchNorm.convert_sync_batchnorm(model)
    n_sync = sum(isinstance(m, nn.SyncBatchNorm) for m in sync.modules())
    assert n_sync == n_bn and n_bn > 0
    assert isinstance(sync.freq_filter.layer1[1], nn.SyncBatchNorm)
    assert isinstance(sync.spat_filter.layer1[1], nn.SyncBatchNorm)
    # timm param_groups_weight_decay rule (engine/forgery_engine.py:152): ndim<=1 or .bias -> no decay
    no_decay = [n for n, p in sync.named_parameters() if p.requires_grad and (p.ndim <= 1 or n.endswith(".bias"))]
    assert "fuse_coef" in no_decay and "freq_filter.layer1.1.weight" in no_decay
    assert "freq_filter.layer1.0.weight" not in no_decay


def test_registry_and_errors():
    from unidefense_b200.model import MODEL, load_model
    assert set(MODEL) == {"UDEB4", "UDR18", "UDR50"}
    with pytest.raises(AssertionError):
        load_model("nope")
    with pytest.raises(ValueError):
        MODEL["UDEB4"]("efficientnet-zz")
    from unidefense_b200.loss import get_loss
    for k in ("aw_triplet", "factorization", "kl_div", "cross_entropy", "bce", "mse"):
        assert isinstance(get_loss(k, "cpu"), nn.Module)


def test_no_cpu_fallback():
    """The product path must fail loudly on CPU tensors instead of silently computing in torch."""
    from unidefense_b200 import ops
    with pytest.raises(RuntimeError):
        ops.in_act(torch.zeros(1, 2, 4, 4), None, None, "relu")
    with pytest.raises(RuntimeError):
        ops.triplet_loss(torch.zeros(4, 8), torch.tensor([0, 0, 1, 1]))
    model = _build("r18").eval()
    with pytest.raises(RuntimeError):
        model(torch.zeros(2, 3, 64, 64))


def test_merge_bn_stats_matches_global_statistics():
    """Cross-rank (mean, M2, count) combination used for SyncBatchNorm equals single-pass statistics."""
    from unidefense_b200.model.modules import merge_bn_stats
    g = torch.Generator().manual_seed(0)
    chunks = [torch.randn(n, 5, generator=g, dtype=torch.float64) * (i + 1) + i for i, n in enumerate([7, 3, 12])]
    means = torch.stack([c.mean(0) for c in chunks])
    m2s = torch.stack([((c - c.mean(0)) ** 2).sum(0) for c in chunks])
    counts = torch.tensor([float(len(c)) for c in chunks], dtype=torch.float64)
    mean, m2, n = merge_bn_stats(means, m2s, counts)
    allx = torch.cat(chunks)
    torch.testing.assert_close(mean, allx.mean(0))
    torch.testing.assert_close(m2, ((allx - allx.mean(0)) ** 2).sum(0))
    assert float(n) == 22


def test_parallel_sync_batchnorm_conversion_cpu():
    """unidefense_b200.parallel.convert_sync_batchnorm: same tree/keys as torch's converter, BatchNorm semantics
    when no process group is initialised (the 2-GPU equivalence with torch.nn.SyncBatchNorm is checked by
    tests/dist_nccl_check.py)."""
    from unidefense_b200.parallel import SyncBatchNorm, convert_sync_batchnorm
    model = _build("r18")
    keys = list(model.state_dict().keys())
    n_bn = sum(isinstance(m, nn.modules.batchnorm._BatchNorm) for m in model.modules())
    conv = convert_sync_batchnorm(model)
    assert list(conv.state_dict().keys()) == keys
    assert sum(isinstance(m, SyncBatchNorm) for m in conv.modules()) == n_bn
    assert isinstance(conv.freq_filter.layer1[1], nn.SyncBatchNorm)          # what model/modules.py keys its sync on
    assert not conv.bottleneck.bias.requires_grad                            # frozen bias survives (shared Parameter)
    assert convert_sync_batchnorm(conv) is conv and sum(isinstance(m, SyncBatchNorm) for m in conv.modules()) == n_bn
    bn = nn.BatchNorm2d(4, momentum=0.01, eps=1e-3)
    sb = convert_sync_batchnorm(nn.Sequential(bn))[0]
    ref = nn.BatchNorm2d(4, momentum=0.01, eps=1e-3)
    ref.load_state_dict(sb.state_dict())
    x = torch.randn(3, 4, 5, 5)
    torch.testing.assert_close(sb(x), ref(x))
    torch.testing.assert_close(sb.running_var, ref.running_var)
    sb.eval(); ref.eval()
    torch.testing.assert_close(sb(x), ref(x))
