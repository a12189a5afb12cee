"""Pins oracle/recon_path.py against fixtures produced by the reference itself
(tests/golden/make_golden.py).  CPU only."""
import torch
import pytest

from oracle import recon_path as O
import procedural as P


def close(a, b, rtol=1e-4, atol=None):
    scale = float(b.abs().max()) if b.numel() else 1.0
    atol = (1e-5 * max(scale, 1e-6)) if atol is None else atol
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol)


def test_interpolate(golden_ops):
    for c in golden_ops["interpolate"]:
        close(O.bilinear_align_corners(c["x"], c["size"]), c["y"], rtol=1e-6, atol=1e-6)


def test_dft_definition():
    for shp in [(2, 3, 6, 5), (1, 2, 7, 8), (1, 1, 19, 12)]:
        x = torch.randn(shp, generator=torch.Generator().manual_seed(1))
        for norm in ("ortho", None):
            a = O.dft_rfft2(x, norm)
            b = torch.fft.rfft2(x, norm=norm)
            close(torch.view_as_real(a), torch.view_as_real(b), rtol=1e-5)


def test_fft_closed_form_grads():
    g = torch.Generator().manual_seed(2)
    for (H, W) in [(6, 6), (5, 7), (12, 12), (8, 5), (19, 20)]:
        for norm in ("ortho", None):
            x = torch.randn(2, 3, H, W, generator=g, dtype=torch.float64, requires_grad=True)
            w = torch.randn(2, 6, H, W // 2 + 1, generator=g, dtype=torch.float64)
            (gx,) = torch.autograd.grad((O.cat_rfft2(x, norm) * w).sum(), x)
            close(O.rfft2_grad_closed_form(w, (H, W), norm), gx, rtol=1e-10, atol=1e-12)
            z = torch.randn(2, 6, H, W // 2 + 1, generator=g, dtype=torch.float64, requires_grad=True)
            gy = torch.randn(2, 3, H, W, generator=g, dtype=torch.float64)
            (gz,) = torch.autograd.grad((O.irfft2_from_cat(z, (H, W), norm) * gy).sum(), z)
            close(O.irfft2_grad_closed_form(gy, norm), gz, rtol=1e-10, atol=1e-12)


def test_in_act(golden_ops):
    for c in golden_ops["in_act"]:
        x = c["x"].clone().requires_grad_()
        g, b = c["gamma"].clone().requires_grad_(), c["beta"].clone().requires_grad_()
        y = O.instance_norm_act(x, g, b, c["act"])
        close(y, c["y"])
        gx, gg, gb = torch.autograd.grad((y * c["gy"]).sum(), [x, g, b])
        close(gx, c["gx"]); close(gg, c["ggamma"]); close(gb, c["gbeta"])
        cx, cg, cb = O.instance_norm_act_backward(c["x"], c["gamma"], c["beta"], c["act"], c["gy"])
        close(cx, c["gx"]); close(cg, c["ggamma"]); close(cb, c["gbeta"])


def test_dynamic_filters(golden_ops):
    for c in golden_ops["dyfi"]:
        sd = c["sd0"]
        x = c["x"].clone().requires_grad_()
        ps = [sd["layer1.0.weight"].clone().requires_grad_(), sd["layer1.1.weight"].clone().requires_grad_(),
              sd["layer1.1.bias"].clone().requires_grad_(), sd["layer2.0.weight"].clone().requires_grad_()]
        mask, out, _ = O.dynamic_filter(x, c["diff"], ps[0], ps[1], ps[2], ps[3], c["act"], True)
        close(mask, c["mask"]); close(out, c["out"])
        gs = torch.autograd.grad((mask * c["gm"]).sum() + (out * c["go"]).sum(), [x] + ps)
        for a, k in zip(gs, ["gx", "gw1", "ggamma", "gbeta", "gw2"]):
            close(a, c[k], rtol=2e-4)
        # running stats after one training step: momentum .1, unbiased variance
        proj = torch.nn.functional.conv2d(c["x"], sd["layer1.0.weight"], padding=sd["layer1.0.weight"].shape[-1] // 2)
        n = proj.numel() // proj.shape[1]
        _, mean, var = O.batch_norm(proj, ps[1], ps[2], None, None, True)
        close(0.9 * sd["layer1.1.running_mean"] + 0.1 * mean.detach(), c["sd1"]["layer1.1.running_mean"])
        close(0.9 * sd["layer1.1.running_var"] + 0.1 * var.detach() * n / (n - 1), c["sd1"]["layer1.1.running_var"])
        sd1 = c["sd1"]
        me, oe, _ = O.dynamic_filter(c["x"], c["diff"], sd1["layer1.0.weight"], sd1["layer1.1.weight"], sd1["layer1.1.bias"],
                                     sd1["layer2.0.weight"], c["act"], False, sd1["layer1.1.running_mean"],
                                     sd1["layer1.1.running_var"])
        close(me, c["mask_eval"]); close(oe, c["out_eval"])


def test_triplet(golden_ops):
    for c in golden_ops["triplet"]:
        f = c["feat"].clone().requires_grad_()
        l = O.aw_triplet_loss(f, c["labels"])
        close(l, c["loss"])
        (g,) = torch.autograd.grad(l, f)
        close(g, c["gfeat"], rtol=1e-3)


def test_factorization(golden_ops):
    for c in golden_ops["factorization"]:
        a = c["a"].clone().requires_grad_()
        l = O.factorization_loss(a, c["b"])
        close(l, c["loss"])
        (g,) = torch.autograd.grad(l, a)
        close(g, c["ga"], rtol=1e-3)


def test_mask_kl(golden_ops):
    for c in golden_ops["mask_kl"]:
        p = c["pred"].clone().requires_grad_()
        l = O.mask_kl_loss(p, c["gt"])
        close(l, c["loss"], atol=1e-8)
        (g,) = torch.autograd.grad(l, p)
        close(g, c["gpred"], rtol=1e-3, atol=1e-8)


def test_perturbations(golden_ops):
    for c in golden_ops["freq_style"]:
        close(O.frequency_style_transfer(c["content"], c["style"], c["lmda"]), c["y"])
    for c in golden_ops["spat_style"]:
        close(O.spatial_style_transfer(c["content"], c["style"], c["lmda"]), c["y"], rtol=1e-6)
    for c in golden_ops["coral"]:
        close(O.coral(c["source"], c["target"]), c["y"])
    for c in golden_ops["blur"]:
        close(O.random_blur(c["x"]), c["y"])
    for c in golden_ops["downscale"]:
        assert torch.equal(O.downscale(c["x"]), c["y"])


def test_sfconv_freq_branch(golden_ops):
    import torch.nn.functional as F
    for c in golden_ops["sfconv"]:
        sd = c["sd"]
        x = c["x"]
        s = c["stride"]
        if c["kind"] == "eff":
            k = sd["weight"].shape[-1]
            ih = x.shape[-1]
            oh = -(-ih // s)
            pad = max((oh - 1) * s + k - ih, 0)
            xp = F.pad(x, (pad // 2, pad - pad // 2, pad // 2, pad - pad // 2))
            spat = F.conv2d(xp, sd["weight"], None, s, 0, 1, x.shape[1])
        else:
            spat = F.conv2d(x, sd["weight"], None, s, 1)
        fr = O.sfconv_freq_branch(x, sd["freq_conv.weight"], spat.shape[-2:], c["norm"])
        co = torch.sigmoid(sd["sf_coef"])
        close((1 - co) * spat + co * fr, c["y"])


def _hot_params(fix):
    arch = fix["arch"]
    p = {}
    for name, shape in O.decoder_param_names(arch):
        p[name] = P.tensor_for(name, shape, salt=3)
    C = O.ATT_DEPTH[arch]
    shapes = {"freq_filter.layer1.0.weight": (2 * C, 2 * C, 1, 1), "freq_filter.layer1.1.weight": (2 * C,),
              "freq_filter.layer1.1.bias": (2 * C,), "freq_filter.layer2.0.weight": (1, 8, 1, 1),
              "spat_filter.layer1.0.weight": (C, C, 3, 3), "spat_filter.layer1.1.weight": (C,),
              "spat_filter.layer1.1.bias": (C,), "spat_filter.layer2.0.weight": (1, 5, 1, 1), "fuse_coef": ()}
    for name, shape in shapes.items():
        p[name] = P.tensor_for(name, shape, salt=3)
    return p


def test_recon_path_against_reference_model(golden_path):
    """decoder + attention + tail + triplet of the oracle vs tensors captured inside the reference
    model classes (UniDefenseModelEb4 / Res18 / Res50.forward) and its autograd gradients."""
    fix = golden_path
    arch = fix["arch"]
    p = {k: v.requires_grad_() for k, v in _hot_params(fix).items()}
    feat = fix["feat"].clone().requires_grad_()
    emb = fix["emb"].clone().requires_grad_()
    x, labels = fix["x"], fix["labels"]
    dec_outs = O.decoder(feat, p, arch)
    for i, d in enumerate(dec_outs, start=1):
        close(d, fix[f"dec_out{i}"], rtol=1e-3)
    att_out, fmask, smask = O.attention(dec_outs[-1].detach(), x, emb, p, O.DECODER_ACT[arch], True)
    close(fmask, fix["freq_mask"], rtol=1e-3); close(smask, fix["spat_mask"], rtol=1e-3)
    close(att_out, fix["att_out"], rtol=1e-3)
    rec, spatial, freq = O.recon_tail(dec_outs[-1], x)
    close(rec, fix["rec"], rtol=1e-3); close(spatial, fix["spatial"]); close(freq, fix["freq"])
    nr = int((labels == 0).sum())
    tri_feats = [fix["feat"].mean(dim=(-2, -1))] + [d.mean(dim=(-2, -1)) for d in dec_outs[:O.TRIPLET_DEC[arch]]]
    for a, b in zip(tri_feats, fix["triplet_feats"]):
        close(a, b, rtol=1e-3)
    tri = sum(O.aw_triplet_loss(f, labels) for f in tri_feats)
    close(tri, fix["triplet_loss"])
    loss = (0.1 * fmask.mean() + 0.1 * smask.mean() + 0.1 * tri + 0.1 * spatial[:nr].mean()
            + 1.0 * freq[:nr].mean() + (att_out * fix["r_att"]).sum())
    close(loss, fix["loss"])
    names = [n for n in fix["param_grads"]]
    gs = torch.autograd.grad(loss, [feat, emb] + [p[n] for n in names])
    close(gs[0], fix["g_feat"], rtol=2e-3, atol=2e-4 * float(fix["g_feat"].abs().max()))
    close(gs[1], fix["g_emb"], rtol=2e-3, atol=2e-4 * float(fix["g_emb"].abs().max()))
    for n, g in zip(names, gs[2:]):
        ref = fix["param_grads"][n]
        assert abs(g.norm().item() - ref["norm"]) <= 2e-3 * ref["norm"] + 1e-7, n
        idx = P.sample_indices(g.numel(), 64, n)
        close(g.flatten()[idx], ref["sample"], rtol=5e-3, atol=2e-3 * float(ref["sample"].abs().max()) + 1e-8)
    # closed-form tail backward vs autograd of the oracle
    gsp = torch.zeros_like(spatial); gsp[:nr] = 0.1 / nr
    gfr = torch.zeros_like(freq); gfr[:nr] = 1.0 / nr
    d = dec_outs[-1].detach().clone().requires_grad_()
    _, s2, f2 = O.recon_tail(d, x)
    (gd,) = torch.autograd.grad((s2 * gsp).sum() + (f2 * gfr).sum(), d)
    close(O.recon_tail_backward_closed_form(d.detach(), x, gsp, gfr), gd, rtol=1e-3, atol=1e-3 * float(gd.abs().max()))


def test_stable_rank_matching_and_spectral_composite():
    """The two oracle functions added for the round-2 kernels: the stable-sort variant equals the reference's on tie-free
    planes (its fixtures) and resolves ties in pixel order; the FFT2 + mask + IFFT2 composite is the identity for an
    all-ones mask and linear in the mask."""
    import torch
    g = torch.Generator().manual_seed(1)
    content = torch.randn(2, 3, 9, 7, generator=g)
    style = torch.randn(2, 3, 9, 7, generator=g)
    lm = torch.rand(2, 1, 1, generator=g) / 2 + 0.5
    torch.testing.assert_close(O.spatial_style_transfer(content, style, lm, stable=True), O.spatial_style_transfer(content, style, lm))
    tied = torch.tensor([2.0, 1.0, 2.0, 1.0]).view(1, 1, 2, 2)
    out = O.spatial_style_transfer(tied, torch.arange(4.0).view(1, 1, 2, 2), torch.zeros(1, 1, 1), stable=True)
    assert out.flatten().tolist() == [2.0, 0.0, 3.0, 1.0]
    x = torch.randn(2, 3, 10, 9, generator=g, dtype=torch.float64)
    ones = torch.ones(2, 10, 5, dtype=torch.float64)
    torch.testing.assert_close(O.spectral_mask_filter(x, ones), x)
    m1, m2 = torch.rand(2, 10, 5, generator=g, dtype=torch.float64), torch.rand(2, 10, 5, generator=g, dtype=torch.float64)
    torch.testing.assert_close(O.spectral_mask_filter(x, m1 + 2 * m2), O.spectral_mask_filter(x, m1) + 2 * O.spectral_mask_filter(x, m2))
