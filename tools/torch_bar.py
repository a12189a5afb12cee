"""The bar: the reference's own op sequences for the hot path, as the STOCK TORCH calls it makes (cuFFT, cuDNN, ATen),
timed on the same GPU next to our kernels (SURVEY.md §8(d) "GPU reference timing").

This is a measurement baseline used by bench.py / the C5 microbench only -- not the oracle (oracle/ is the checker)
and never on the product path.  Every function cites the reference lines whose torch calls it issues.
"""
import torch
import torch.nn.functional as F


class _Swish(torch.autograd.Function):
    """model/efficientnet/utils.py:66-77 SwishImplementation (custom autograd: 2 extra passes each way)."""

    @staticmethod
    def forward(ctx, i):
        ctx.save_for_backward(i)
        return i * torch.sigmoid(i)

    @staticmethod
    def backward(ctx, g):
        i = ctx.saved_tensors[0]
        s = torch.sigmoid(i)
        return g * (s * (1 + i * (1 - s)))


def in_act(x, gamma, beta, act):
    """nn.InstanceNorm2d(affine) + MemoryEfficientSwish / nn.ReLU (model/unidefense.py:61-98, :286-305)."""
    y = F.instance_norm(x, None, None, gamma, beta, True, 0.1, 1e-5)
    return _Swish.apply(y) if act == "swish" else F.relu(y)


def recon_tail(dec, x, norm="ortho"):
    """model/unidefense.py:244-253."""
    rec = F.interpolate(dec, size=x.shape[-2:], mode="bilinear", align_corners=True)
    spatial = torch.abs(rec - x).mean([-3, -2, -1])
    rf = torch.fft.rfft2(rec, dim=(-2, -1), norm=norm)
    rf = torch.cat([rf.real, rf.imag], dim=1)
    xf = torch.fft.rfft2(x, dim=(-2, -1), norm=norm)
    xf = torch.cat([xf.real, xf.imag], dim=1)
    tmp = torch.abs(rf - xf)
    re, im = tmp.tensor_split(2, dim=1)
    freq = (re + im).mean([-3, -2, -1])
    return rec, spatial, freq


def _dyfi(x, diff, w1, gamma, beta, w2, act, rm, rv):
    """model/modules.py:91-105 / :120-134 (training mode)."""
    proj = F.conv2d(x, w1, None, 1, w1.shape[-1] // 2)
    proj = F.batch_norm(proj, rm, rv, gamma, beta, True, 0.1, 1e-5)
    proj = _Swish.apply(proj) if act == "swish" else F.relu(proj)
    avg = torch.mean(proj, dim=1, keepdim=True)
    mx, _ = torch.max(proj, dim=1, keepdim=True)
    mask = torch.sigmoid(F.conv2d(torch.cat([avg, mx, diff], dim=1), w2))
    return mask, mask * x


def attention(pred, x, emb, p, act, norm="ortho"):
    """model/unidefense.py:125-157 (dropout inactive)."""
    size = emb.shape[-2:]
    pd = F.interpolate(pred, size=size, mode="bilinear", align_corners=True)
    xs = F.interpolate(x, size=size, mode="bilinear", align_corners=True)
    pf = torch.fft.rfft2(pd, dim=(-2, -1), norm=norm)
    pf = torch.cat([pf.real, pf.imag], dim=1)
    xf = torch.fft.rfft2(xs, dim=(-2, -1), norm=norm)
    xf = torch.cat([xf.real, xf.imag], dim=1)
    ef = torch.fft.rfft2(emb, dim=(-2, -1), norm=norm)
    ef = torch.cat([ef.real, ef.imag], dim=1)
    fmask, fout = _dyfi(ef, torch.abs(pf - xf), p["fw1"], p["fg"], p["fb"], p["fw2"], act, p["frm"], p["frv"])
    re, im = torch.tensor_split(fout, 2, dim=1)
    ff = torch.fft.irfft2(torch.complex(re, im), s=size, dim=(-2, -1), norm=norm)
    smask, sout = _dyfi(emb, torch.abs(pd - xs), p["sw1"], p["sg"], p["sb"], p["sw2"], act, p["srm"], p["srv"])
    c = torch.sigmoid(p["coef"])
    return (1 - c) * sout + c * ff + emb.clone(), fmask, smask


def attention_params(C, dev, seed=3):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g).to(dev)         # noqa: E731
    return {"fw1": (r(2 * C, 2 * C, 1, 1) / (2 * C) ** 0.5).requires_grad_(), "fg": (r(2 * C) * 0.1 + 1).requires_grad_(),
            "fb": (r(2 * C) * 0.1).requires_grad_(), "fw2": (r(1, 8, 1, 1) * 0.5).requires_grad_(),
            "frm": torch.zeros(2 * C, device=dev), "frv": torch.ones(2 * C, device=dev),
            "sw1": (r(C, C, 3, 3) / (9 * C) ** 0.5).requires_grad_(), "sg": (r(C) * 0.1 + 1).requires_grad_(),
            "sb": (r(C) * 0.1).requires_grad_(), "sw2": (r(1, 5, 1, 1) * 0.5).requires_grad_(),
            "srm": torch.zeros(C, device=dev), "srv": torch.ones(C, device=dev),
            "coef": torch.tensor(0.3, device=dev).requires_grad_()}


def triplet(feat, labels):
    """loss/triplet_loss.py:16-82 (host sync in torch.where included: it is what the reference executes)."""
    m = feat.size(0)
    xx = torch.pow(feat, 2).sum(1, keepdim=True).expand(m, m)
    dist = xx + xx.t()
    dist = dist - 2 * torch.matmul(feat, feat.t())
    dist = dist.clamp(min=1e-12).sqrt()
    N = m
    n_real = torch.where(1 - labels)[0].shape[0]
    ne = ~torch.eye(N, device=labels.device, dtype=torch.bool)
    is_pos = labels.expand(N, N).eq(labels.expand(N, N).t()) & ne
    is_neg = labels.expand(N, N).ne(labels.expand(N, N).t())
    ap = dist[:n_real][is_pos[:n_real]].reshape(n_real, -1)
    an = dist[:n_real][is_neg[:n_real]].reshape(n_real, -1)
    eap, ean = torch.exp(ap), torch.exp(-an)
    wp = eap / (eap.sum(1, keepdim=True) + 1e-12)
    wn = ean / (ean.sum(1, keepdim=True) + 1e-12)
    fwp, fwn = torch.sum(wp * ap, dim=1), torch.sum(wn * an, dim=1)
    return F.soft_margin_loss(fwn - fwp, torch.ones_like(fwn))


def freq_style_transfer(content, style, lmda):
    """model/modules.py:43-54."""
    H, W = content.shape[-2:]
    lm = lmda.reshape(-1, 1, 1, 1)
    fa = torch.fft.rfft2(content, dim=(-2, -1), norm="ortho")
    am, ap = torch.abs(fa), torch.angle(fa)
    bm = torch.abs(torch.fft.rfft2(style, dim=(-2, -1), norm="ortho"))
    rec = (lm * am + (1.0 - lm) * bm) * torch.exp(1j * ap)
    return torch.fft.irfft2(rec, s=(H, W), dim=(-2, -1), norm="ortho")


def spectral_mask_l1(x, mask, target):
    """C5 composite (BASELINE.json configs[4]): FFT2 -> half-spectrum mask -> IFFT2 -> L1 loss against a target."""
    H, W = x.shape[-2:]
    y = torch.fft.irfft2(torch.fft.rfft2(x, dim=(-2, -1), norm="ortho") * mask, s=(H, W), dim=(-2, -1), norm="ortho")
    return y, torch.abs(y - target).mean([-3, -2, -1])


def time_cuda(fn, iters, flush, warmup=3):
    """Mean / best ms of fn() between CUDA events on the current stream, L2 flushed before every iteration."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    tot, best = 0.0, 1e30
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1)
        tot += t
        best = min(best, t)
    return tot / iters, best
