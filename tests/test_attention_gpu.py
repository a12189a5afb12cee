"""a4/a7 parity: bilinear resize, error maps, feature-map rFFT2/irFFT2 (+autograd), fuse."""
import pytest
import torch

from oracle import recon_path as O

pytestmark = pytest.mark.gpu


def close(a, b, rtol=1e-4, atol=None):
    b = b.detach()
    scale = float(b.abs().max()) if b.numel() else 1.0
    atol = (1e-5 * max(scale, 1e-6)) if atol is None else atol
    torch.testing.assert_close(a.detach().cpu().to(b.dtype), b, rtol=rtol, atol=atol)


def test_bilinear_reference_fixture(golden_ops):
    from unidefense_b200 import ops
    for c in golden_ops["interpolate"]:
        close(ops.bilinear_ac(c["x"].cuda(), c["size"]), c["y"], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("shape,size", [((2, 3, 5, 6), (11, 12)), ((1, 2, 38, 38), (12, 12)), ((2, 3, 380, 380), (12, 12)),
                                        ((1, 3, 7, 7), (7, 7)), ((1, 1, 4, 4), (1, 1)), ((2, 2, 1, 1), (3, 3))])
def test_bilinear_fwd_bwd(shape, size):
    from unidefense_b200 import ops
    g = torch.Generator().manual_seed(0)
    x = torch.randn(shape, generator=g)
    gy = torch.randn(*shape[:2], *size, generator=g)
    xc = x.cuda().requires_grad_()
    y = ops.bilinear_ac(xc, size)
    (y * gy.cuda()).sum().backward()
    x32 = x.clone().requires_grad_()
    y32 = torch.nn.functional.interpolate(x32, size=size, mode="bilinear", align_corners=True)
    (y32 * gy).sum().backward()
    close(y, y32, rtol=1e-5, atol=2e-6)
    close(xc.grad, x32.grad, rtol=1e-4)


SIZES = [(12, 12), (8, 8), (16, 16), (24, 24), (7, 7), (5, 6), (9, 4), (19, 19), (1, 1), (2, 3), (48, 48), (64, 64), (3, 64)]
# two-kernel path through the workspace (a side > 64): mixed radix (95 = 5*19, 380, 150, 75), Bluestein (74 = 2*37,
# 62 = 2*31, 124, 67 prime, 248 = 8*31), odd / even widths, degenerate sides, non-square
SIZES += [(95, 95), (65, 65), (380, 380), (150, 75), (74, 74), (62, 124), (67, 3), (3, 67), (1, 100), (100, 1), (248, 96),
          (128, 80)]


@pytest.mark.parametrize("hw", SIZES)
@pytest.mark.parametrize("norm", ["ortho", None])
def test_rfft2_irfft2_fwd_bwd(hw, norm):
    from unidefense_b200 import ops
    h, w = hw
    g = torch.Generator().manual_seed(h * 100 + w)
    N, C = (2, 5) if h * w < 50000 else (1, 2)
    x = torch.randn(N, C, h, w, generator=g)
    wgt = torch.randn(N, 2 * C, h, w // 2 + 1, generator=g)
    xc = x.cuda().requires_grad_()
    xf = ops.rfft2_cat(xc, norm)
    (xf * wgt.cuda()).sum().backward()
    x64 = x.double().requires_grad_()
    xf64 = O.cat_rfft2(x64, norm)
    (xf64 * wgt.double()).sum().backward()
    close(xf, xf64)
    close(xc.grad, x64.grad)

    z = torch.randn(N, 2 * C, h, w // 2 + 1, generator=g)
    gy = torch.randn(N, C, h, w, generator=g)
    zc = z.cuda().requires_grad_()
    y = ops.irfft2_cat(zc, (h, w), norm)
    (y * gy.cuda()).sum().backward()
    z64 = z.double().requires_grad_()
    y64 = O.irfft2_from_cat(z64, (h, w), norm)
    (y64 * gy.double()).sum().backward()
    close(y, y64)
    close(zc.grad, z64.grad)


def test_rfft2_config_shapes_roundtrip():
    """Full config sizes (UDEB4 N=32: [32,272,12,12]; UDR50 N=64: [64,2048,8,8]): irfft2(rfft2(x)) == x."""
    from unidefense_b200 import ops
    for shape in [(32, 272, 12, 12), (64, 2048, 8, 8), (4, 512, 24, 24)]:
        x = torch.randn(shape, device="cuda")
        y = ops.irfft2_cat(ops.rfft2_cat(x), shape[-2:])
        torch.testing.assert_close(y, x, rtol=1e-4, atol=2e-5)
        # Parseval (ortho): sum |X|^2 over the Hermitian-completed spectrum == sum x^2
        xf = ops.rfft2_cat(x)
        w = shape[-1]
        wt = torch.full((w // 2 + 1,), 2.0, device="cuda")
        wt[0] = 1.0
        if w % 2 == 0:
            wt[-1] = 1.0
        torch.testing.assert_close((xf.double() ** 2 * wt).sum(), (x.double() ** 2).sum(), rtol=1e-5, atol=0)


def test_unsupported_and_empty():
    from unidefense_b200 import ops
    with pytest.raises(RuntimeError):
        ops.rfft2_cat(torch.zeros(1, 1, 1025, 8, device="cuda"))       # > UD_FFT_MAX_N
    assert ops.rfft2_cat(torch.zeros(0, 4, 8, 8, device="cuda")).shape == (0, 8, 8, 5)
    with pytest.raises(ValueError):
        ops.irfft2_cat(torch.zeros(1, 4, 8, 4, device="cuda"), (8, 8))


@pytest.mark.parametrize("case", [((2, 3, 64, 64), (2, 3, 128, 128), (4, 4)), ((2, 3, 192, 192), (2, 3, 380, 380), (12, 12)),
                                  ((3, 3, 38, 38), (3, 3, 76, 76), (5, 5)), ((1, 3, 128, 128), (1, 3, 256, 256), (8, 8)),
                                  ((2, 3, 190, 190), (2, 3, 380, 380), (24, 24)), ((2, 3, 20, 30), (2, 3, 41, 59), (7, 9))])
def test_attn_prep(case):
    from unidefense_b200 import ops
    ps, xs, size = case
    g = torch.Generator().manual_seed(4)
    pred = torch.tanh(torch.randn(ps, generator=g))
    x = torch.rand(xs, generator=g) * 2 - 1
    sd, fd = ops.attn_prep(pred.cuda(), x.cuda(), size)
    sd32, fd32 = O.attention_prep(pred, x, size)
    # 4-tap bilinear of fp32 data: coordinates in fp32 like ATen
    close(sd, sd32, rtol=1e-4, atol=2e-6)
    close(fd, fd32, rtol=1e-4, atol=1e-5)
    sd2, fd2 = ops.attn_prep(pred.cuda(), x.cuda(), size, None)
    _, fd32n = O.attention_prep(pred, x, size, None)
    close(fd2, fd32n, rtol=1e-4, atol=1e-5 * float(fd32n.abs().max()))


@pytest.mark.parametrize("shape", [(2, 5, 4, 3), (3, 272, 12, 12), (2, 64, 8, 8), (1, 7, 24, 24), (2, 3, 1, 1)])
@pytest.mark.parametrize("with_res", [False, True])
def test_attn_fuse(shape, with_res):
    from unidefense_b200 import ops
    N, C, h, w = shape
    g = torch.Generator().manual_seed(7)
    emb, ff, res = (torch.randn(shape, generator=g) for _ in range(3))
    smask = torch.rand(N, 1, h, w, generator=g)
    coef = torch.tensor(0.3)
    gout = torch.randn(shape, generator=g)
    ts = [t.cuda().requires_grad_() for t in (emb, smask, ff, res, coef)]
    out = ops.attn_fuse(ts[0], ts[1], ts[2], ts[3] if with_res else None, ts[4])
    (out * gout.cuda()).sum().backward()
    t64 = [t.double().requires_grad_() for t in (emb, smask, ff, res, coef)]
    s = torch.sigmoid(t64[4])
    o64 = (1 - s) * t64[1] * t64[0] + s * t64[2] + (t64[3] if with_res else t64[0])
    (o64 * gout.double()).sum().backward()
    close(out, o64)
    for i, name in enumerate(["emb", "smask", "ff", "res", "coef"]):
        if name == "res" and not with_res:
            assert ts[i].grad is None
            continue
        close(ts[i].grad, t64[i].grad, rtol=2e-4, atol=1e-5 * float(t64[i].grad.abs().max()) + 1e-7)


@pytest.mark.parametrize("shape", [(2, 3, 380, 380), (2, 3, 224, 224), (1, 2, 62, 80), (3, 3, 24, 24)])
def test_spectral_mask_filter_composite(shape):
    """The C5 composite irfft2(mask * rfft2(x)) (generic transforms, mask folded into the inverse's load) vs torch.fft."""
    from unidefense_b200 import ops
    N, C, H, W = shape
    g = torch.Generator().manual_seed(H + W)
    x = torch.randn(shape, generator=g)
    mask = torch.rand(N, H, W // 2 + 1, generator=g)
    y = ops.spectral_mask_filter(x.cuda(), mask.cuda())
    close(y, O.spectral_mask_filter(x.double(), mask.double()))
    with pytest.raises(ValueError):
        ops.spectral_mask_filter(x.cuda(), mask[:, :-1].cuda())
