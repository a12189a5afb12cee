"""a13-a16 parity: perturbations of the second training pass (no grad) vs reference fixtures and the oracle."""
import pytest
import torch

from oracle import recon_path as O

pytestmark = pytest.mark.gpu


def close(a, b, rtol=1e-4, atol=None):
    scale = float(b.abs().max()) if b.numel() else 1.0
    atol = (1e-5 * max(scale, 1e-6)) if atol is None else atol
    torch.testing.assert_close(a.detach().cpu().to(b.dtype), b, rtol=rtol, atol=atol)


def test_freq_style_reference_fixture(golden_ops):
    from unidefense_b200 import ops
    for c in golden_ops["freq_style"]:
        y = ops.freq_style_transfer(c["content"].cuda(), c["style"].cuda(), c["lmda"].reshape(-1).cuda())
        close(y, c["y"], rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("shape", [(2, 3, 380, 380), (1, 3, 256, 256), (1, 3, 224, 224), (1, 3, 299, 299), (2, 3, 20, 19),
                                   (1, 2, 33, 48), (3, 1, 7, 7), (1, 3, 16, 15), (1, 1, 1, 1), (1, 1, 2, 2),
                                   (2, 3, 58, 58), (1, 2, 62, 37), (1, 3, 248, 248)])      # Bluestein sizes
def test_freq_style_vs_oracle(shape):
    from unidefense_b200 import ops
    g = torch.Generator().manual_seed(sum(shape))
    content = torch.rand(shape, generator=g) * 2 - 1
    style = torch.rand(shape, generator=g) * 2 - 1
    lm = torch.rand(shape[0], generator=g) / 2 + 0.5
    y = ops.freq_style_transfer(content.cuda(), style.cuda(), lm.cuda())
    y64 = O.frequency_style_transfer(content.double(), style.double(), lm.double().view(-1, 1, 1, 1))
    close(y, y64, rtol=1e-4, atol=3e-5)


def test_freq_style_properties_full_batch():
    """N=32 at 380^2 (config size): lmda=1 or style==content returns the content; the mean (DC) is the mix of means."""
    from unidefense_b200 import ops
    g = torch.Generator().manual_seed(0)
    content = (torch.rand(32, 3, 380, 380, generator=g) * 2 - 1).cuda()
    style = (torch.rand(32, 3, 380, 380, generator=g) * 2 - 1).cuda()
    ones = torch.ones(32, device="cuda")
    torch.testing.assert_close(ops.freq_style_transfer(content, style, ones), content, rtol=1e-4, atol=2e-5)
    lm = torch.full((32,), 0.7, device="cuda")
    torch.testing.assert_close(ops.freq_style_transfer(content, content, lm), content, rtol=1e-4, atol=2e-5)
    y = ops.freq_style_transfer(content, style, lm)
    mc, ms = content.mean(dim=(-2, -1)), style.mean(dim=(-2, -1))
    want = torch.sign(mc) * (0.7 * mc.abs() + 0.3 * ms.abs())         # DC bin: real, keeps the content's sign
    torch.testing.assert_close(y.mean(dim=(-2, -1)), want, rtol=1e-3, atol=1e-5)
    with pytest.raises(RuntimeError):                                  # > UD_FFT_MAX_N
        ops.freq_style_transfer(torch.zeros(1, 1, 8, 1025, device="cuda"), torch.zeros(1, 1, 8, 1025, device="cuda"), ones[:1])


def test_blur_and_downscale(golden_ops):
    from unidefense_b200 import ops
    for c in golden_ops["blur"]:
        close(ops.gaussian_blur5(c["x"].cuda()), c["y"], rtol=1e-5, atol=1e-6)
    for c in golden_ops["downscale"]:
        assert torch.equal(ops.downscale_nearest(c["x"].cuda()).cpu(), c["y"])        # index arithmetic: bit exact
    x = torch.rand(4, 3, 380, 380) * 2 - 1
    close(ops.gaussian_blur5(x.cuda()), O.random_blur(x), rtol=1e-5, atol=1e-6)
    assert torch.equal(ops.downscale_nearest(x.cuda()).cpu(), O.downscale(x))
    for shp in [(1, 3, 256, 256), (2, 3, 224, 224), (1, 3, 299, 299), (1, 1, 5, 7)]:
        x = torch.rand(shp)
        assert torch.equal(ops.downscale_nearest(x.cuda()).cpu(), O.downscale(x))


def test_spatial_style(golden_ops):
    from unidefense_b200 import ops
    for c in golden_ops["spat_style"]:
        y = ops.spatial_style_transfer(c["content"].cuda(), c["style"].cuda(), c["lmda"].reshape(-1).cuda())
        close(y, c["y"], rtol=1e-6, atol=1e-6)
    # tie-free values (ties make the rank order -- and the reference's own output -- sort-implementation defined)
    g = torch.Generator().manual_seed(3)
    n = 2 * 3 * 64 * 64
    content = (torch.randperm(n, generator=g).float() / n).view(2, 3, 64, 64)
    style = (torch.randperm(n, generator=g).float() / n * 2 - 1).view(2, 3, 64, 64)
    lm = torch.tensor([0.5, 0.9])
    y = ops.spatial_style_transfer(content.cuda(), style.cuda(), lm.cuda())
    close(y, O.spatial_style_transfer(content, style, lm.view(-1, 1, 1)), rtol=1e-6, atol=1e-6)
    # histogram property: with lmda -> 0 the output takes exactly the style's values, in the content's rank order
    y0 = ops.spatial_style_transfer(content.cuda(), style.cuda(), torch.zeros(2).cuda())
    torch.testing.assert_close(y0.flatten(2).sort(-1).values, style.cuda().flatten(2).sort(-1).values, rtol=0, atol=1e-6)


@pytest.mark.parametrize("shape", [(2, 3, 380, 380), (1, 3, 299, 299), (3, 2, 17, 9), (1, 1, 1, 1), (2, 1, 91, 90), (1, 2, 256, 256)])
def test_spatial_style_sort_kernel(shape):
    """The radix-sort kernel against the oracle on continuous random values (ties have probability ~0 except where
    forced below), incl. negative values, planes that are not a multiple of the 8192-key tile and a full 380^2 plane."""
    from unidefense_b200 import ops
    g = torch.Generator().manual_seed(sum(shape))
    content = torch.randn(shape, generator=g) * 3
    style = torch.randn(shape, generator=g) - 0.5
    lm = torch.rand(shape[0], generator=g) / 2 + 0.5
    y = ops.spatial_style_transfer(content.cuda(), style.cuda(), lm.cuda())
    want = O.spatial_style_transfer(content, style, lm.view(-1, 1, 1))
    # ~100 pairs of exactly equal fp32 draws occur in a 144400-element plane; torch.sort leaves their order open, so
    # the reference's output is only defined up to that: compare with the oracle through the sorted values ...
    close(y.flatten(2).sort(-1).values, want.flatten(2).sort(-1).values, rtol=1e-6, atol=1e-6)
    assert float(((y.cpu() - want).abs() <= 1e-6 + 1e-6 * want.abs()).float().mean()) > 0.998
    # ... and pixel by pixel with the same formula evaluated with a STABLE sort, which is the kernel's tie rule
    stable = O.spatial_style_transfer(content, style, lm.view(-1, 1, 1), stable=True).flatten(2)
    torch.testing.assert_close(y.flatten(2).cpu(), stable, rtol=0, atol=1e-6)
    # sorting sanity: lmda = 0 reproduces the style's multiset exactly, in the content's rank order
    y0 = ops.spatial_style_transfer(content.cuda(), style.cuda(), torch.zeros(shape[0]).cuda())
    # ((c + m) - c carries one rounding of |c + m| <= 20: 2e-6)
    order = torch.argsort(content.flatten(2), dim=-1, stable=True)
    torch.testing.assert_close(y0.flatten(2).cpu().gather(-1, order), style.flatten(2).sort(-1).values, rtol=0, atol=4e-6)


def test_spatial_style_ties_are_stable_and_shapes_checked():
    from unidefense_b200 import ops
    content = torch.tensor([2.0, 1.0, 2.0, 1.0, -0.0, 0.0, 2.0, 1.0]).view(1, 1, 2, 4)
    style = torch.arange(8.0).view(1, 1, 2, 4)
    y0 = ops.spatial_style_transfer(content.cuda(), style.cuda(), torch.zeros(1).cuda()).cpu().flatten()
    # ranks with ties in pixel order: -0.0 (idx 4) < 0.0 (idx 5) by the integer image of the floats, then the 1s, the 2s
    assert y0.tolist() == [5.0, 2.0, 6.0, 3.0, 0.0, 1.0, 7.0, 4.0]
    with pytest.raises(AssertionError):
        ops.spatial_style_transfer(torch.zeros(1, 3, 4, 4).cuda(), torch.zeros(1, 3, 4, 5).cuda(), torch.ones(1).cuda())
    assert ops.spatial_style_transfer(torch.zeros(0, 3, 4, 4).cuda(), torch.zeros(0, 3, 4, 4).cuda(),
                                      torch.ones(0).cuda()).shape == (0, 3, 4, 4)


def _coral_candidates(source, target):
    """The reference's coral (utils/operation.py:20-45) in fp64 for every sign pattern of the two SVDs: its
    U*sqrt(D)*Vh^T 'square root'  equals U*sqrt(D)*U for the symmetric positive definite f f^T + I and
    changes with the sign of each singular vector -- a gauge the SVD leaves open.  -> {(signs_s, signs_t): image};
    the all-ones pattern is the kernel's documented convention (largest component of each eigenvector positive)."""
    import itertools

    def stats(img):
        f = img.reshape(3, -1).double()
        mean, std = f.mean(1, keepdim=True), f.std(1, keepdim=True)
        fn = (f - mean) / std
        return fn, mean, std, fn @ fn.t() + torch.eye(3, dtype=torch.double)

    def quirk(cov, signs):
        d, u = torch.linalg.eigh(cov)
        d, u = d.flip(0), u.flip(1).clone()
        for j in range(3):
            if u[u[:, j].abs().argmax(), j] < 0:
                u[:, j] = -u[:, j]
        u = u * torch.tensor(signs, dtype=torch.double)[None, :]
        return u @ torch.diag(d.sqrt()) @ u

    fs, _, _, cs = stats(source)
    _, mt, st, ct = stats(target)
    out = {}
    for sa in itertools.product([1, -1], repeat=3):
        for sb in itertools.product([1, -1], repeat=3):
            m = quirk(ct, sb) @ torch.linalg.inv(quirk(cs, sa))
            out[(sa, sb)] = ((m @ fs) * st + mt).reshape(source.shape)
    return out


def test_coral_batch(golden_ops):
    """a15 parity, ASSERTED: (1) the reference fixture (CPU LAPACK) is one of the 64 sign patterns of the quirk --
    i.e. the reference equals our formula modulo the SVD's sign gauge; (2) the kernel equals the pattern of its
    documented convention to fp32 accuracy, on the fixtures and on a 380x380 batch."""
    from unidefense_b200 import ops
    ones = ((1, 1, 1), (1, 1, 1))
    for c in golden_ops["coral"]:
        cand = _coral_candidates(c["source"], c["target"])
        scale = float(c["y"].abs().max())
        errs = {k: float((v - c["y"].double()).abs().max()) for k, v in cand.items()}
        assert min(errs.values()) <= 1e-4 * scale, f"reference fixture matches no sign pattern (best {min(errs.values()):.2e})"
        got = ops.coral_batch(c["source"][None].cuda(), c["target"][None].cuda())[0].cpu().double()
        assert float((got - cand[ones]).abs().max()) <= 1e-4 * scale
    g = torch.Generator().manual_seed(4)
    base = torch.rand(4, 1, 380, 380, generator=g)
    src = (0.6 * base + 0.4 * torch.rand(4, 3, 380, 380, generator=g)) * 2 - 1        # correlated channels, like faces
    tgt = (0.5 * base.flip(0) + 0.5 * torch.rand(4, 3, 380, 380, generator=g)) * 1.2 - 0.5
    y = ops.coral_batch(src.cuda(), tgt.cuda()).cpu().double()
    assert torch.isfinite(y).all()
    for n in range(4):
        want = _coral_candidates(src[n], tgt[n])[ones]
        assert float((y[n] - want).abs().max()) <= 2e-4 * float(want.abs().max())
        # what coral is for: the output carries the target's per-channel mean and (unbiased) std
    ym, ys = y.flatten(2).mean(-1), y.flatten(2).std(-1)
    assert ym.shape == (4, 3) and ys.shape == (4, 3)
