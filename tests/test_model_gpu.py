"""Model-level parity on the GPU: (1) the hot path driven with the boundary tensors captured inside the
reference model classes (path_*.pt); (2) the whole drop-in model, backbone included, against the whole
reference model with name-regenerated weights (full_*.pt)."""
import os

import pytest
import torch
import torch.nn.functional as F

import procedural as P

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CTOR = {"eb4": ("UDEB4", dict(extractor="efficientnet-b4", num_classes=2, drop_rate=0.0, drop_connect_rate=0.0)),
        "r18": ("UDR18", dict(num_classes=2, drop_rate=0.0)),
        "r50": ("UDR50", dict(extractor="resnet50", num_classes=2, drop_rate=0.0))}


def close(a, b, rtol=1e-3, atol=None):
    b = b.detach()
    scale = float(b.abs().max()) if b.numel() else 1.0
    atol = (1e-4 * max(scale, 1e-6)) if atol is None else atol
    torch.testing.assert_close(a.detach().cpu().to(b.dtype), b, rtol=rtol, atol=atol)


def _build(arch, salt, hot_only):
    from unidefense_b200.model import load_model
    name, kw = CTOR[arch]
    model = load_model(name)(**kw)
    P.fill_state_dict_(model, prefix_filter=P.is_hot if hot_only else None, salt=salt)
    return model.cuda().train()


@pytest.fixture()
def no_dropout(monkeypatch):
    monkeypatch.setattr(F, "dropout", lambda t, p=0.5, training=True, inplace=False: t * 1.0)


def test_hot_path_against_reference_model(golden_path, no_dropout):
    from unidefense_b200 import ops
    fix = golden_path
    arch = fix["arch"]
    model = _build(arch, 3, True)
    feat = fix["feat"].cuda().requires_grad_()
    emb = fix["emb"].cuda().requires_grad_()
    x, labels = fix["x"].cuda(), fix["labels"].cuda()
    nblocks = 2 if arch == "r18" else 3
    ntri = 2 if arch == "eb4" else 1
    y, dec_outs, tris = feat, [], []
    for i in range(1, nblocks + 1):
        blk = getattr(model, f"dec_block{i}")
        if i <= ntri:
            y, t = blk.forward_with_mean(y)
            tris.append(t)
        else:
            y = blk(y)
        dec_outs.append(y)
        close(y, fix[f"dec_out{i}"])
    att = model.attention(dec_outs[-1].detach(), x, emb)
    close(att["freq_mask"], fix["freq_mask"]); close(att["spat_mask"], fix["spat_mask"]); close(att["out"], fix["att_out"])
    rec, spatial, freq = ops.recon_tail(dec_outs[-1], x)
    close(rec, fix["rec"]); close(spatial, fix["spatial"], rtol=1e-4); close(freq, fix["freq"], rtol=1e-4)
    # x_b4 / ext_feat enters the triplet list BEFORE the decoder-input dropout node whose output the fixture
    # captured as `feat`, so g_feat is the decoder-path gradient only (tests/golden/make_golden.py:make_path)
    tri_feats = [fix["feat"].cuda().mean(dim=(-2, -1))] + tris
    for a, b in zip(tri_feats, fix["triplet_feats"]):
        close(a, b)
    tri = sum(ops.triplet_loss(f, labels) for f in tri_feats)
    close(tri, fix["triplet_loss"], rtol=1e-4)
    nr = int((fix["labels"] == 0).sum())
    loss = (0.1 * att["freq_mask"].mean() + 0.1 * att["spat_mask"].mean() + 0.1 * tri + 0.1 * spatial[:nr].mean()
            + 1.0 * freq[:nr].mean() + (att["out"] * fix["r_att"].cuda()).sum())
    close(loss, fix["loss"], rtol=1e-4)
    names = list(fix["param_grads"])
    params = dict(model.named_parameters())
    gs = torch.autograd.grad(loss, [feat, emb] + [params[n] for n in names])
    # Input gradients pass through |.| kinks (sign(d), sign(Re/Im dF)) and, for r50 at R=64, a BatchNorm over only
    # 16 values per channel: a handful of near-zero bins taking the other sign (cuDNN vs MKL summation order) moves
    # individual elements by ~1 % of the maximum.  Element-wise closeness holds for most elements; the hard
    # criterion is the relative error in norm.
    for got, want, what in ((gs[0], fix["g_feat"], "g_feat"), (gs[1], fix["g_emb"], "g_emb")):
        err = float((got.detach().cpu() - want).norm() / want.norm())
        assert err < 1e-2, f"{what}: relative L2 error {err:.3e}"
        frac_bad = float(((got.detach().cpu() - want).abs() > 2e-3 * want.abs() + 2e-2 * float(want.abs().max())).float().mean())
        assert frac_bad == 0.0, f"{what}: {frac_bad:.3%} of elements off by more than 2 % of the maximum"
    for n, g in zip(names, gs[2:]):
        ref = fix["param_grads"][n]
        assert abs(g.norm().item() - ref["norm"]) <= 2e-3 * ref["norm"] + 1e-7, n
        idx = P.sample_indices(g.numel(), 64, n)
        # sampled entries: same |.|-kink conditioning as the input gradients above (2 % of the sample's maximum)
        close(g.flatten()[idx.cuda()], ref["sample"], rtol=5e-3, atol=2e-2 * float(ref["sample"].abs().max()) + 1e-8)
    sd = model.state_dict()
    for k, v in fix["bn_after"].items():
        close(sd[k], v, rtol=1e-4)


@pytest.mark.parametrize("arch", ["eb4", "r18", "r50"])
def test_full_model_against_reference(arch, no_dropout):
    from unidefense_b200 import ops
    fix = torch.load(os.path.join(GOLDEN, f"full_{arch}.pt"), weights_only=False)
    model = _build(arch, 5, False)
    x = P.tensor_for(f"in:full_x_{arch}", (fix["N"], 3, fix["R"], fix["R"]), "unit").cuda()
    labels = fix["labels"].cuda()
    out = model(x)
    ld = out["loss_dict"]
    assert set(out) == {"cls_out", "rec", "loss_dict"}
    assert set(ld) == {"factorization", "triplet", "freq_mask", "spat_mask", "spatial", "freq"}
    noise = fix["noise"]       # the reference's own spread under a 1-ulp input perturbation (make_golden.make_full)

    def near(a, b, nz, what):
        b = b.detach()
        tol = 1e-4 * max(float(b.abs().max()), 1e-6) + 8.0 * nz
        err = float((a.detach().cpu() - b).abs().max())
        assert err <= tol, f"{what}: max abs err {err:.3e} > {tol:.3e} (reference noise {nz:.3e})"

    near(ld["spatial"], fix["spatial"], noise["spatial"], "spatial")
    near(ld["freq"], fix["freq"], noise["freq"], "freq")
    near(ld["freq_mask"], fix["freq_mask"], noise["freq_mask"], "freq_mask")
    near(ld["spat_mask"], fix["spat_mask"], noise["spat_mask"], "spat_mask")
    near(out["rec"][:, :, ::7, ::5], fix["rec_sample"], noise["rec_sample"], "rec")
    for i, (a, b) in enumerate(zip(ld["triplet"], fix["triplet_feats"])):
        near(a, b, noise["triplet_feats"][i], f"triplet[{i}]")
    near(ld["factorization"], fix["factorization"], noise["factorization"], "factorization")
    near(out["cls_out"], fix["cls_out"], noise["cls_out"], "cls_out")
    nr = fix["N"] // 2
    tri = sum(ops.triplet_loss(f, labels) for f in ld["triplet"])
    loss = (F.cross_entropy(out["cls_out"], labels) + 0.1 * ld["freq_mask"].mean() + 0.1 * ld["spat_mask"].mean()
            + 0.1 * tri + 0.1 * ld["spatial"][:nr].mean() + 1.0 * ld["freq"][:nr].mean())
    near(loss, fix["loss"], noise["loss"], "loss")
    loss.backward()
    # gradient norms: a structural check (a missing term is an O(1) error).  One perturbation run is a crude
    # estimate of the reference's conditioning and cannot see the model's kinks: a channel-argmax of the dynamic
    # filters that flips under 1e-6 jitter of the decoder output moves fuse_coef / spat_filter.layer2 by 4.3 % / 3.9 %
    # (both states measured, profiles/r02_determinism_udr18.txt), so the floor is 6 %; parameters whose true gradient
    # is zero (a bias feeding another BatchNorm) hold only rounding residue and are skipped by magnitude.
    bad = []
    gmax = max(v["norm"] for v in fix["param_grads"].values() if v is not None)
    for n, p in model.named_parameters():
        ref = fix["param_grads"].get(n)
        if not p.requires_grad:
            continue
        assert ref is not None, f"{n}: reference has no gradient entry"
        assert p.grad is not None, f"{n}: no gradient (DDP find_unused_parameters=False would hang)"
        gn = p.grad.norm().item()
        if ref["norm"] < 1e-5 * gmax:
            continue
        tol = (6e-2 + 8.0 * noise["grad_norm_rel"][n]) * ref["norm"]
        if abs(gn - ref["norm"]) > tol:
            bad.append((n, gn, ref["norm"], noise["grad_norm_rel"][n]))
    assert not bad, bad[:8]                  # (name, ours, reference, reference noise)
    sd = model.state_dict()
    for k, v in fix["bn_after"].items():
        near(sd[k], v, noise["bn_after"][k], k)


def test_eval_mode_and_dtypes():
    """validate()/test() path (engine/forgery_engine.py:343-350): eval mode, no grad, softmax over cls_out."""
    model = _build("r18", 5, False).eval()
    x = (torch.rand(3, 3, 64, 64, device="cuda") * 2 - 1)
    with torch.no_grad():
        out = model(x)
    assert out["cls_out"].shape == (3, 2) and out["rec"].shape == x.shape
    assert out["loss_dict"]["freq_mask"].shape == (3, 1, 4, 3) and out["loss_dict"]["spat_mask"].shape == (3, 1, 4, 4)
    assert torch.isfinite(torch.softmax(out["cls_out"], 1)).all()
    out2 = model(x)
    torch.testing.assert_close(out2["cls_out"], out["cls_out"])      # deterministic


# bf16 statement for the BENCHMARKED configuration (bench.py: bf16 autocast for the stock-torch backbone and the dense
# convolutions, channels_last backbone, SFConv transforms as bf16 DFT-by-GEMM, TF32 tcgen05 projections, fp32 hot-path
# kernels).  bf16 carries 8 mantissa bits (2^-9 = 2e-3 per rounding) and the models normalise with train-mode batch /
# instance statistics, so single elements can move a lot (a pixel of a low-variance InstanceNorm plane, a logit behind
# a 4-sample BatchNorm1d) while every aggregate stays close.  The statement is therefore in RELATIVE L2 error
# ||a - b|| / ||b|| per tensor (plus the absolute error of the scalar losses), stated twice:
#   (1) against the fp32 REFERENCE fixtures (4 samples: the noisiest possible batch statistics);
#   (2) against this repo's own fp32 path -- which the tests above pin to the reference at 1e-4 -- on 16 samples, where
#       the batch statistics are what training sees.
def _rel2(a, b):
    a, b = a.detach().float().cpu().flatten(), b.detach().float().cpu().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def _bench_config(model):
    for part in ("backbone", "extractor", "emb_block1", "emb_block2"):      # bench.py --channels-last backbone
        if hasattr(model, part):
            getattr(model, part).to(memory_format=torch.channels_last)
    return model


def _bf16_forward(model, x):
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = True                                   # torch default, as in bench.py
    try:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return model(x)
    finally:
        torch.backends.cudnn.allow_tf32 = old


def _pass1_loss(out, labels, nr):
    from unidefense_b200 import ops
    ld = out["loss_dict"]
    tri = sum(ops.triplet_loss(f.float(), labels) for f in ld["triplet"])
    return (ops.cross_entropy(out["cls_out"].float(), labels) + 0.1 * ld["freq_mask"].mean() + 0.1 * ld["spat_mask"].mean()
            + 0.1 * tri + 0.1 * ld["spatial"][:nr].mean() + 1.0 * ld["freq"][:nr].mean())


# Measured on a B200 (tools/bf16_attribution.py, N=16 at 128^2; relative L2 vs the fp32 path):
#   UDR18:  rec 0.067  logits 0.065  freq_mask 0.031  spat_mask 0.011  triplet 0.002  losses 2e-4
#   UDEB4:  rec 0.231  logits 0.346  freq_mask 0.138  spat_mask 0.030  triplet 0.079  losses 8e-4
# and the SAME numbers (+-0.01) with the DFT-by-GEMM SFConv switched off (cuFFT in fp32), with an NCHW backbone, and with
# 3xTF32 instead of TF32 projections: the deviation is bf16 autocast of the stock 32-block EfficientNet at random init
# (swish + BatchNorm eps 1e-3 on 16 samples), not something this repo's kernels or layouts add.  Budgets = 1.5x measured.
# relative-L2 budgets per arch: (vs reference fixtures, 4 samples), (vs own fp32 path, 16 samples)
BF16_BUDGET = {
    "r18": {"rec": (0.10, 0.10), "freq_mask": (0.05, 0.05), "spat_mask": (0.03, 0.03), "triplet": (0.01, 0.01),
            "cls_out": (0.10, 0.10), "spatial": (0.003, 0.003), "freq": (0.003, 0.003), "loss": (0.01, 0.01)},
    "eb4": {"rec": (0.35, 0.35), "freq_mask": (0.22, 0.21), "spat_mask": (0.05, 0.05), "triplet": (0.12, 0.12),
            "cls_out": (0.50, 0.52), "spatial": (0.003, 0.003), "freq": (0.003, 0.003), "loss": (0.06, 0.05)}}


def _bf16_report(pairs, which, arch, label):
    bad, seen = [], []
    for what, a, b in pairs:
        err, tol = _rel2(a, b), BF16_BUDGET[arch][what.split("[")[0]][which]
        seen.append(f"{what} {err:.2e}/{tol:.0e}")
        if not err <= tol:
            bad.append(f"{what}: relative L2 error {err:.3e} > {tol:.1e}")
    print(f"bf16 bench configuration vs {label} ({arch}), relative L2 error / budget: " + "; ".join(seen))
    assert not bad, f"bf16 bench configuration vs {label}: " + "; ".join(bad) + " | all: " + "; ".join(seen)


@pytest.mark.parametrize("arch", ["eb4", "r18"])
def test_bench_configuration_bf16_against_fp32_reference(arch, no_dropout):
    fix = torch.load(os.path.join(GOLDEN, f"full_{arch}.pt"), weights_only=False)
    model = _bench_config(_build(arch, 5, False))
    x = P.tensor_for(f"in:full_x_{arch}", (fix["N"], 3, fix["R"], fix["R"]), "unit").cuda()
    labels = fix["labels"].cuda()
    out = _bf16_forward(model, x)
    ld = out["loss_dict"]
    loss = _pass1_loss(out, labels, fix["N"] // 2)
    pairs = [("spatial", ld["spatial"], fix["spatial"]), ("freq", ld["freq"], fix["freq"]),
             ("freq_mask", ld["freq_mask"], fix["freq_mask"]), ("spat_mask", ld["spat_mask"], fix["spat_mask"]),
             ("rec", out["rec"][:, :, ::7, ::5], fix["rec_sample"]), ("cls_out", out["cls_out"], fix["cls_out"]),
             ("loss", loss, fix["loss"])]
    pairs += [(f"triplet[{i}]", a, b) for i, (a, b) in enumerate(zip(ld["triplet"], fix["triplet_feats"]))]
    _bf16_report(pairs, 0, arch, "the fp32 reference fixtures (N=4)")
    loss.backward()
    assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in model.parameters() if p.requires_grad)


@pytest.mark.parametrize("arch,res", [("eb4", 128), ("r18", 128)])
def test_bench_configuration_bf16_against_own_fp32_path(arch, res, no_dropout):
    n = 16
    g = torch.Generator().manual_seed(77)
    x = (torch.rand(n, 3, res, res, generator=g) * 2 - 1).cuda()
    labels = torch.tensor([0] * (n // 2) + [1] * (n // 2)).cuda()
    ref_model = _build(arch, 5, False)
    ref = ref_model(x)
    ref_loss = _pass1_loss(ref, labels, n // 2)
    model = _bench_config(_build(arch, 5, False))
    out = _bf16_forward(model, x)
    loss = _pass1_loss(out, labels, n // 2)
    a, b = out["loss_dict"], ref["loss_dict"]
    pairs = [("spatial", a["spatial"], b["spatial"]), ("freq", a["freq"], b["freq"]), ("freq_mask", a["freq_mask"], b["freq_mask"]),
             ("spat_mask", a["spat_mask"], b["spat_mask"]), ("rec", out["rec"], ref["rec"]), ("cls_out", out["cls_out"], ref["cls_out"]),
             ("loss", loss, ref_loss)]
    pairs += [(f"triplet[{i}]", u, v) for i, (u, v) in enumerate(zip(a["triplet"], b["triplet"]))]
    _bf16_report(pairs, 1, arch, "this repo's fp32 path (N=16)")
    # gradients: direction and size of the flat parameter gradient
    loss.backward()
    ref_loss.backward()
    ga = torch.cat([p.grad.float().flatten() for p in model.parameters() if p.grad is not None])
    gb = torch.cat([p.grad.float().flatten() for p in ref_model.parameters() if p.grad is not None])
    cos = float(torch.dot(ga, gb) / (ga.norm() * gb.norm()))
    ratio = float(ga.norm() / gb.norm())
    print(f"bf16 vs fp32 flat parameter gradient ({arch}): cosine {cos:.4f}, norm ratio {ratio:.4f}")
    # (reported; the hard statement is about the forward quantities above: at random init the logits of the 32-block
    # EfficientNet move by a third under bf16, and the parameter gradient follows them)
    assert cos > (0.9 if arch == "r18" else 0.3) and 0.5 < ratio < 2.0, (cos, ratio)
