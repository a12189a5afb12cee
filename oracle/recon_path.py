"""CPU oracle for the UniDefense dual-space reconstruction hot path.

TEST INFRASTRUCTURE ONLY.  This file is a plain, functional restatement (torch CPU tensor
ops, dtype-generic so fp64 works) of the algorithms on the hot path of the reference
(VISION-SJTU/UniDefense, read-only at /root/reference).  It is the checker for the CUDA
kernels in unidefense_b200/csrc; it is never imported by the product package.  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import it.

Parity pin: the reference ships NO tests, golden vectors or KATs, so
this oracle is pinned against the reference itself, executed live in the build container:
tests/golden/make_golden.py imports /root/reference (with timm/matplotlib stubs), runs the
reference modules on seeded inputs and commits the inputs/outputs as fixtures under
tests/golden/*.pt; tests/test_oracle_golden.py checks every function below against them.

Third-party arithmetic on the path (not under /root/reference): PyTorch (README pins
1.12.1; this image has 2.11.0) -- torch.fft.rfft2/irfft2 and their autograd formulas,
F.interpolate, instance/batch norm, sort, linalg.svd; torchvision gaussian_blur.  Their
published algorithms are restated here explicitly (DFT definition, align_corners bilinear
index arithmetic, biased-variance normalisation, 5x5 sigma=1.1 reflect-padded blur) and
torch.fft is used only as a fast evaluator of the same DFT (checked against `dft_rfft2`).

Each function cites the reference file:line it follows.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------
# bilinear resize, align_corners=True          (model/unidefense.py:16 `interpolate`)
# ----------------------------------------------------------------------------------------
def _ac_axis(in_size: int, out_size: int, dtype, device):
    """Source index / weights along one axis (ATen area_pixel_compute_source_index with
    align_corners=True: scale=(in-1)/(out-1) in the op's accumulate type, src=scale*dst)."""
    acc = torch.float32 if dtype in (torch.float32, torch.bfloat16, torch.float16) else dtype
    if out_size > 1:
        scale = torch.tensor(in_size - 1, dtype=acc) / torch.tensor(out_size - 1, dtype=acc)
    else:
        scale = torch.tensor(0, dtype=acc)
    src = scale * torch.arange(out_size, dtype=acc)
    i0 = src.to(torch.int64)                      # trunc == floor (src >= 0)
    i1 = i0 + (i0 < in_size - 1).to(torch.int64)
    l1 = (src - i0.to(acc)).to(dtype)
    l0 = (1.0 - l1).to(dtype)
    return i0.to(device), i1.to(device), l0.to(device), l1.to(device)


def bilinear_align_corners(x: Tensor, size: Sequence[int]) -> Tensor:
    """F.interpolate(x, size, mode='bilinear', align_corners=True) restated."""
    H, W = int(size[0]), int(size[1])
    h, w = x.shape[-2:]
    y0, y1, ly0, ly1 = _ac_axis(h, H, x.dtype, x.device)
    x0, x1, lx0, lx1 = _ac_axis(w, W, x.dtype, x.device)
    top = x.index_select(-2, y0)
    bot = x.index_select(-2, y1)
    a, b = top.index_select(-1, x0), top.index_select(-1, x1)
    c, d = bot.index_select(-1, x0), bot.index_select(-1, x1)
    ly0 = ly0[:, None]
    ly1 = ly1[:, None]
    return ly0 * (lx0 * a + lx1 * b) + ly1 * (lx0 * c + lx1 * d)


# ----------------------------------------------------------------------------------------
# real 2-D DFT in the reference's channel-planar cat([re, im], dim=1) layout
# ----------------------------------------------------------------------------------------
def dft_rfft2(x: Tensor, norm: Optional[str] = "ortho") -> Tensor:
    """Definition-level rfft2 (O(n^2) per line), complex output [..., H, W//2+1].
    X[j,k] = s * sum_{r,c} x[r,c] exp(-2 pi i (j r / H + k c / W)); s = 1/sqrt(HW) for
    'ortho', 1 for None/'backward'."""
    H, W = x.shape[-2:]
    Wh = W // 2 + 1
    dt = torch.float64
    kc = torch.outer(torch.arange(W, dtype=dt), torch.arange(Wh, dtype=dt)) * (2 * math.pi / W)
    jr = torch.outer(torch.arange(H, dtype=dt), torch.arange(H, dtype=dt)) * (2 * math.pi / H)
    Fw = torch.complex(torch.cos(kc), -torch.sin(kc))          # [W, Wh]
    Fh = torch.complex(torch.cos(jr), -torch.sin(jr))          # [H, H]
    xc = x.to(dt).to(torch.complex128)
    y = torch.matmul(Fh, torch.matmul(xc, Fw))
    if norm == "ortho":
        y = y / math.sqrt(H * W)
    return y.to(torch.complex64 if x.dtype == torch.float32 else torch.complex128)


def cat_rfft2(x: Tensor, norm: Optional[str] = "ortho") -> Tensor:
    """rfft2 then cat([re, im], dim=1)          (model/unidefense.py:130-136, :246-250)."""
    f = torch.fft.rfft2(x, norm=norm)
    return torch.cat([f.real, f.imag], dim=1)


def irfft2_from_cat(xf: Tensor, size: Sequence[int], norm: Optional[str] = "ortho") -> Tensor:
    """complex(*tensor_split(xf, 2, dim=1)) then irfft2(s=size)  (model/unidefense.py:142-145)."""
    re, im = torch.tensor_split(xf, 2, dim=1)
    return torch.fft.irfft2(torch.complex(re.contiguous(), im.contiguous()),
                            s=tuple(int(s) for s in size), norm=norm)


def irfft2_grad_closed_form(gy: Tensor, norm: Optional[str] = "ortho") -> Tensor:
    """Backward of irfft2_from_cat w.r.t. xf (SURVEY App. B.3 == torch fft_c2r_backward):
    G = rfft2(gy) with the *inverse's* normalisation, interior columns doubled; DC/Nyquist
    columns keep their (non-zero) imaginary parts.  Returns cat([re, im], dim=1)."""
    H, W = gy.shape[-2:]
    Wh = W // 2 + 1
    g = torch.fft.rfft2(gy, norm="ortho" if norm == "ortho" else "forward")
    last = Wh - 1 if W % 2 == 0 else Wh
    scale = torch.ones(Wh, dtype=gy.dtype, device=gy.device)
    scale[1:last] = 2.0
    g = g * scale
    return torch.cat([g.real, g.imag], dim=1)


def rfft2_grad_closed_form(gf: Tensor, size: Sequence[int], norm: Optional[str] = "ortho") -> Tensor:
    """Backward of cat_rfft2 w.r.t. its real input (SURVEY App. B.2 == torch
    fft_r2c_backward): zero-pad the half spectrum to full width (no Hermitian mirroring),
    inverse complex FFT, keep the real part."""
    H, W = int(size[0]), int(size[1])
    re, im = torch.tensor_split(gf, 2, dim=1)
    Wh = re.shape[-1]
    S = torch.zeros(*re.shape[:-1], W, dtype=torch.complex64 if gf.dtype == torch.float32
                    else torch.complex128, device=gf.device)
    S[..., :Wh] = torch.complex(re.contiguous(), im.contiguous())
    # adjoint of the forward transform: 'ortho' -> ortho inverse; None -> unnormalised inverse
    g = torch.fft.ifft2(S, norm="ortho" if norm == "ortho" else "forward")
    return g.real


# ----------------------------------------------------------------------------------------
# a1: reconstruction-loss tail                 (model/unidefense.py:244-253, :423-433, :618-628)
# ----------------------------------------------------------------------------------------
def recon_tail(dec: Tensor, x: Tensor, norm: Optional[str] = "ortho") -> Tuple[Tensor, Tensor, Tensor]:
    """rec = bilinear(dec -> x.shape); spatial[n] = mean|rec-x|; freq[n] = mean(|Re dF|+|Im dF|)
    with dF = rfft2(rec) - rfft2(x) -- the reference's two-FFT formulation, verbatim order."""
    rec = bilinear_align_corners(dec, x.shape[-2:])
    spatial = torch.abs(rec - x).mean(dim=[-3, -2, -1])
    rec_f = cat_rfft2(rec, norm)
    x_f = cat_rfft2(x, norm)
    tmp = torch.abs(rec_f - x_f)
    t_re, t_im = tmp.tensor_split(2, dim=1)
    freq = (t_re + t_im).mean(dim=[-3, -2, -1])
    return rec, spatial, freq


def recon_tail_backward_closed_form(dec: Tensor, x: Tensor, g_spatial: Tensor, g_freq: Tensor,
                                    norm: Optional[str] = "ortho") -> Tensor:
    """d(sum_n g_spatial[n]*spatial[n] + g_freq[n]*freq[n]) / d dec, by the closed forms of
    SURVEY App. B.1/B.2/B.6 (sign spectra -> zero-padded inverse FFT -> transposed bilinear)."""
    N, C, h, w = dec.shape
    H, W = x.shape[-2:]
    Wh = W // 2 + 1
    rec = bilinear_align_corners(dec, (H, W))
    d = rec - x
    D = cat_rfft2(d, norm)                                     # linearity (App. B.1)
    gs = (g_spatial / (C * H * W)).view(N, 1, 1, 1)
    gf = (g_freq / (C * H * Wh)).view(N, 1, 1, 1)
    g_rec = gs * torch.sign(d) + rfft2_grad_closed_form(gf * torch.sign(D), (H, W), norm)
    # transposed bilinear (scatter-add along each axis)
    y0, y1, ly0, ly1 = _ac_axis(h, H, dec.dtype, dec.device)
    x0, x1, lx0, lx1 = _ac_axis(w, W, dec.dtype, dec.device)
    tmp = torch.zeros(N, C, H, w, dtype=dec.dtype)
    tmp.index_add_(-1, x0, g_rec * lx0)
    tmp.index_add_(-1, x1, g_rec * lx1)
    out = torch.zeros(N, C, h, w, dtype=dec.dtype)
    out.index_add_(-2, y0, tmp * ly0[:, None])
    out.index_add_(-2, y1, tmp * ly1[:, None])
    return out


# ----------------------------------------------------------------------------------------
# a2: decoder epilogues                        (model/unidefense.py:54-56,:61-101; efficientnet/utils.py:66-82)
# ----------------------------------------------------------------------------------------
def swish(x: Tensor) -> Tensor:
    return x * torch.sigmoid(x)


def swish_backward(x: Tensor, g: Tensor) -> Tensor:
    """SwishImplementation.backward (efficientnet/utils.py:73-77)."""
    s = torch.sigmoid(x)
    return g * (s * (1 + x * (1 - s)))


def activation(x: Tensor, act: str) -> Tensor:
    if act == "swish":
        return swish(x)
    if act == "relu":
        return torch.relu(x)
    if act == "none":
        return x
    raise ValueError(act)


def instance_norm_act(x: Tensor, gamma: Optional[Tensor], beta: Optional[Tensor], act: str,
                      eps: float = 1e-5) -> Tensor:
    """nn.InstanceNorm2d(affine, no running stats) + activation: per-(n,c) plane biased
    variance.                                   (model/unidefense.py:61-62 etc.)"""
    mu = x.mean(dim=(-2, -1), keepdim=True)
    var = ((x - mu) ** 2).mean(dim=(-2, -1), keepdim=True)
    y = (x - mu) / torch.sqrt(var + eps)
    if gamma is not None:
        y = y * gamma.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)
    return activation(y, act)


def instance_norm_act_backward(x: Tensor, gamma: Optional[Tensor], beta: Optional[Tensor], act: str,
                               gy: Tensor, eps: float = 1e-5):
    """Closed-form backward of instance_norm_act -> (gx, ggamma, gbeta)."""
    mu = x.mean(dim=(-2, -1), keepdim=True)
    var = ((x - mu) ** 2).mean(dim=(-2, -1), keepdim=True)
    rstd = 1.0 / torch.sqrt(var + eps)
    xh = (x - mu) * rstd
    g_ = gamma.view(1, -1, 1, 1) if gamma is not None else 1.0
    b_ = beta.view(1, -1, 1, 1) if beta is not None else 0.0
    z = xh * g_ + b_
    if act == "swish":
        gz = swish_backward(z, gy)
    elif act == "relu":
        gz = gy * (z > 0).to(gy.dtype)
    else:
        gz = gy
    ggamma = (gz * xh).sum(dim=(0, 2, 3))
    gbeta = gz.sum(dim=(0, 2, 3))
    gxh = gz * g_
    m1 = gxh.mean(dim=(-2, -1), keepdim=True)
    m2 = (gxh * xh).mean(dim=(-2, -1), keepdim=True)
    gx = rstd * (gxh - m1 - xh * m2)
    return gx, ggamma, gbeta


# decoder layout per architecture: list of blocks, each a list of (kind, c_in, c_out)
# kind: 'c' = Conv3x3 s1 p1 + IN + act ; 't' = ConvT3x3 s2 p1 op1 + IN + act ; 'o' = Conv3x3 + tanh
DECODER_SPEC = {
    # model/unidefense.py:59-102
    "eb4": [[("c", 160, 80), ("t", 80, 80), ("c", 80, 80)],
            [("c", 80, 40), ("t", 40, 40), ("c", 40, 40)],
            [("c", 40, 20), ("t", 20, 20), ("c", 20, 20), ("o", 20, 3)]],
    # model/unidefense.py:284-308 (mid_depth 448)
    "r18": [[("c", 448, 128), ("t", 128, 128), ("c", 128, 128)],
            [("c", 128, 64), ("t", 64, 64), ("c", 64, 32), ("o", 32, 3)]],
    # model/unidefense.py:464-500 (mid_depth 1024)
    "r50": [[("c", 1024, 256), ("t", 256, 256), ("c", 256, 256)],
            [("c", 256, 128), ("t", 128, 128), ("c", 128, 128)],
            [("c", 128, 64), ("t", 64, 64), ("c", 64, 32), ("o", 32, 3)]],
}
DECODER_ACT = {"eb4": "swish", "r18": "relu", "r50": "relu"}
ATT_DEPTH = {"eb4": 272, "r18": 512, "r50": 2048}
# how many decoder block outputs join the backbone feature as triplet features (unidefense.py:232-236, :411-414, :606-609)
TRIPLET_DEC = {"eb4": 2, "r18": 1, "r50": 1}


def decoder_param_names(arch: str) -> List[Tuple[str, Tuple[int, ...]]]:
    """state_dict names/shapes of the decoder in the reference's nn.Sequential indexing
    (convs at 0,3,6,(9); InstanceNorm weight/bias at 1,4,7)  (SURVEY App. A.4)."""
    out = []
    for bi, block in enumerate(DECODER_SPEC[arch], start=1):
        idx = 0
        for kind, ci, co in block:
            if kind == "t":
                out.append((f"dec_block{bi}.{idx}.weight", (ci, co, 3, 3)))
            else:
                out.append((f"dec_block{bi}.{idx}.weight", (co, ci, 3, 3)))
            if kind in ("c", "t"):
                out.append((f"dec_block{bi}.{idx + 1}.weight", (co,)))
                out.append((f"dec_block{bi}.{idx + 1}.bias", (co,)))
                idx += 3
            else:
                idx += 2
    return out


def decoder(feat: Tensor, params: Dict[str, Tensor], arch: str) -> List[Tensor]:
    """Runs dec_block1..k on (already dropped-out) features; returns every block output.
    (model/unidefense.py:213-217, :393-395, :586-590)"""
    act = DECODER_ACT[arch]
    outs = []
    y = feat
    for bi, block in enumerate(DECODER_SPEC[arch], start=1):
        idx = 0
        for kind, ci, co in block:
            w = params[f"dec_block{bi}.{idx}.weight"]
            if kind == "t":
                y = F.conv_transpose2d(y, w, None, stride=2, padding=1, output_padding=1)
            else:
                y = F.conv2d(y, w, None, stride=1, padding=1)
            if kind in ("c", "t"):
                y = instance_norm_act(y, params.get(f"dec_block{bi}.{idx + 1}.weight"),
                                      params.get(f"dec_block{bi}.{idx + 1}.bias"), act)
                idx += 3
            else:
                y = torch.tanh(y)
                idx += 2
        outs.append(y)
    return outs


# ----------------------------------------------------------------------------------------
# a4-a7: attention()                           (model/unidefense.py:125-157; model/modules.py:79-134)
# ----------------------------------------------------------------------------------------
def attention_prep(pred: Tensor, x: Tensor, size: Sequence[int], norm: Optional[str] = "ortho"):
    """Error maps that guide the filters (no grad): spat_diff [N,3,h,w], freq_diff [N,6,h,wh]
    (model/unidefense.py:126-134,:148) -- two-FFT formulation as in the reference."""
    p = bilinear_align_corners(pred, size)
    xs = bilinear_align_corners(x, size)
    freq_diff = torch.abs(cat_rfft2(p, norm) - cat_rfft2(xs, norm))
    spat_diff = torch.abs(p - xs)
    return spat_diff, freq_diff


def batch_norm(x: Tensor, gamma: Tensor, beta: Tensor, running_mean: Optional[Tensor],
               running_var: Optional[Tensor], training: bool, eps: float = 1e-5):
    """nn.BatchNorm{1,2}d forward; returns (y, batch_mean, batch_var_biased).  In training the
    caller updates running stats with momentum 0.1 and the *unbiased* variance."""
    dims = [0] + list(range(2, x.dim()))
    shape = [1, -1] + [1] * (x.dim() - 2)
    if training:
        mean = x.mean(dim=dims)
        var = ((x - mean.view(shape)) ** 2).mean(dim=dims)
    else:
        mean, var = running_mean, running_var
    y = (x - mean.view(shape)) / torch.sqrt(var.view(shape) + eps)
    y = y * gamma.view(shape) + beta.view(shape)
    return y, mean, var


def dynamic_filter(x: Tensor, diff: Tensor, w1: Tensor, bn_gamma: Tensor, bn_beta: Tensor, w2: Tensor,
                   act: str, training: bool = True, running_mean: Optional[Tensor] = None,
                   running_var: Optional[Tensor] = None):
    """FrequencyDynamicFilter.forward (w1 [2C,2C,1,1], w2 [1,8,1,1]; modules.py:91-105) and
    SpatialDynamicFilter.forward (w1 [C,C,3,3], w2 [1,5,1,1]; modules.py:120-134).
    Returns (mask, filtered, proj)."""
    pad = w1.shape[-1] // 2
    proj = F.conv2d(x, w1, None, stride=1, padding=pad)
    proj, _, _ = batch_norm(proj, bn_gamma, bn_beta, running_mean, running_var, training)
    proj = activation(proj, act)
    pre = torch.cat([proj.mean(dim=1, keepdim=True), proj.max(dim=1, keepdim=True).values, diff], dim=1)
    mask = torch.sigmoid(F.conv2d(pre, w2))
    return mask, mask * x, proj


def attention(pred: Tensor, x: Tensor, emb: Tensor, p: Dict[str, Tensor], act: str,
              training: bool = True, dropped_emb: Optional[Tensor] = None,
              norm: Optional[str] = "ortho"):
    """attention() of the three models (model/unidefense.py:125-157).  `dropped_emb` stands for
    self.dropout(embedding.clone()) (RNG lives with the caller); defaults to emb.
    `p` holds freq_filter.* / spat_filter.* / fuse_coef under their state_dict names."""
    size = emb.shape[-2:]
    spat_diff, freq_diff = attention_prep(pred, x, size, norm)
    emb_freq = cat_rfft2(emb, norm)
    fmask, ffilt, _ = dynamic_filter(
        emb_freq, freq_diff, p["freq_filter.layer1.0.weight"], p["freq_filter.layer1.1.weight"],
        p["freq_filter.layer1.1.bias"], p["freq_filter.layer2.0.weight"], act, training,
        p.get("freq_filter.layer1.1.running_mean"), p.get("freq_filter.layer1.1.running_var"))
    freq_filtered = irfft2_from_cat(ffilt, size, norm)
    smask, sfilt, _ = dynamic_filter(
        emb, spat_diff, p["spat_filter.layer1.0.weight"], p["spat_filter.layer1.1.weight"],
        p["spat_filter.layer1.1.bias"], p["spat_filter.layer2.0.weight"], act, training,
        p.get("spat_filter.layer1.1.running_mean"), p.get("spat_filter.layer1.1.running_var"))
    c = torch.sigmoid(p["fuse_coef"])
    out = (1.0 - c) * sfilt + c * freq_filtered
    out = out + (emb if dropped_emb is None else dropped_emb)
    return out, fmask, smask


# ----------------------------------------------------------------------------------------
# a9: asymmetrical weighted triplet loss       (loss/triplet_loss.py:16-82)
# ----------------------------------------------------------------------------------------
def euclidean_dist(x: Tensor) -> Tensor:
    """loss/triplet_loss.py:16-30 with y = x."""
    sq = (x * x).sum(1, keepdim=True)
    dist = sq + sq.t() - 2.0 * (x @ x.t())
    return dist.clamp(min=1e-12).sqrt()


def aw_triplet_loss(feat: Tensor, labels: Tensor) -> Tensor:
    """Anchors = rows with label 0 (must come first); positives = other label-0 rows,
    negatives = label-1 rows; softmax(+d) / softmax(-d) weights with eps 1e-12 in the
    denominators; SoftMarginLoss(wn - wp, 1) = mean log(1+exp(-(wn-wp)))."""
    eps = 1e-12
    N = feat.shape[0]
    n_real = int((labels == 0).sum())
    dist = euclidean_dist(feat)
    losses = []
    for i in range(n_real):
        pos = [j for j in range(N) if j != i and int(labels[j]) == int(labels[i])]
        neg = [j for j in range(N) if int(labels[j]) != int(labels[i])]
        dp = dist[i, pos]
        dn = dist[i, neg]
        ep = torch.exp(dp)
        en = torch.exp(-dn)
        wp = (ep / (ep.sum() + eps) * dp).sum()
        wn = (en / (en.sum() + eps) * dn).sum()
        losses.append(torch.log1p(torch.exp(-(wn - wp))))
    return torch.stack(losses).mean()


# ----------------------------------------------------------------------------------------
# a10: factorization ("calibration") loss      (loss/calib_loss.py:17-28)
# ----------------------------------------------------------------------------------------
def factorization_loss(emb_a: Tensor, emb_b: Tensor, off_diag_weight: float = 0.005, eps: float = 1e-6) -> Tensor:
    n, f = emb_a.shape
    a = (emb_a - emb_a.mean(0)) / (emb_a.std(0) + eps)          # unbiased std
    b = (emb_b - emb_b.mean(0)) / (emb_b.std(0) + eps)
    c = a.t() @ b / n
    diag = torch.diagonal(c)
    on = ((diag - 1.0) ** 2).mean()
    off = ((c ** 2).sum() - (diag ** 2).sum()) / (f * (f - 1))
    return on + off_diag_weight * off


# ----------------------------------------------------------------------------------------
# a11: mask losses                             (engine/abstract_engine.py:215-228, :331-357)
# ----------------------------------------------------------------------------------------
def mask_kl_loss(mask_pred: Tensor, mask_gt: Tensor) -> Tensor:
    """KLDivLoss(batchmean, log_target=True)(log_softmax(pred.flat), log_softmax(gt.flat))."""
    n = mask_pred.shape[0]
    lp = torch.log_softmax(mask_pred.reshape(n, -1), dim=-1)
    lg = torch.log_softmax(mask_gt.reshape(n, -1), dim=-1)
    return (torch.exp(lg) * (lg - lp)).sum() / n


# ----------------------------------------------------------------------------------------
# a13-a16: perturbations (no grad)             (model/modules.py:7-76; utils/operation.py:15-45)
# ----------------------------------------------------------------------------------------
def frequency_style_transfer(content: Tensor, style: Tensor, lmda: Tensor) -> Tensor:
    """modules.py:36-55 with the CPU-RNG draw `lmda` ([B,1,1,1], already in [0.5,1)) passed in."""
    H, W = content.shape[-2:]
    fa = torch.fft.rfft2(content, norm="ortho")
    fb = torch.fft.rfft2(style, norm="ortho")
    am, ap = torch.abs(fa), torch.angle(fa)
    bm = torch.abs(fb)
    mix = (lmda * am + (1.0 - lmda) * bm) * torch.exp(1j * ap)
    return torch.fft.irfft2(mix, s=(H, W), norm="ortho")


def spatial_style_transfer(content: Tensor, style: Tensor, lmda: Tensor, stable: bool = False) -> Tensor:
    """modules.py:59-76 (exact histogram matching); lmda [B,1,1].  The reference sorts with torch.sort's default, which
    leaves the order of EQUAL content values open; `stable=True` fixes it to pixel order -- the CUDA kernel's tie rule
    (csrc/ud_style_sort.cu) -- so that planes with repeated values have one defined answer."""
    B, C, H, W = content.shape
    cf = content.reshape(B, C, -1)
    _, idx = torch.sort(cf, dim=-1, stable=stable)
    vs, _ = torch.sort(style.reshape(B, C, -1), dim=-1)
    inv = idx.argsort(-1)
    out = cf + (1 - lmda) * vs.gather(-1, inv) - (1 - lmda) * cf
    return out.view(B, C, H, W)


def spectral_mask_filter(x: Tensor, mask: Tensor, norm="ortho") -> Tensor:
    """irfft2(mask * rfft2(x)): the FFT2 + mask + IFFT2 composite of BASELINE.json configs[4]; the mask [N,H,W/2+1] is
    shared by the channels of a sample as the dynamic filter's is (model/unidefense.py:135-145)."""
    H, W = x.shape[-2:]
    return torch.fft.irfft2(torch.fft.rfft2(x, norm=norm) * mask.unsqueeze(1), s=(H, W), norm=norm)


def coral_stats(img: Tensor):
    """utils/operation.py:6-12,:24-27: per-channel mean / unbiased std, normalised pixels and
    the un-normalised covariance f f^T + I."""
    f = img.reshape(3, -1)
    mean = f.mean(dim=-1, keepdim=True)
    std = f.std(dim=-1, keepdim=True)
    fn = (f - mean) / std
    cov = fn @ fn.t() + torch.eye(3, dtype=img.dtype)
    return fn, mean, std, cov


def mat_sqrt_quirk(m: Tensor) -> Tensor:
    """utils/operation.py:15-17.  torch.linalg.svd returns Vh, the reference applies .t() to it as
    if it were V, i.e. computes U diag(sqrt(D)) Vh^T -- NOT the matrix square root (App. D)."""
    U, D, Vh = torch.linalg.svd(m)
    return U @ torch.diag(D.sqrt()) @ Vh.t()


def coral(source: Tensor, target: Tensor) -> Tensor:
    """utils/operation.py:20-45 for one [3,H,W] pair."""
    s_n, _, _, s_cov = coral_stats(source)
    _, t_mean, t_std, t_cov = coral_stats(target)
    m = mat_sqrt_quirk(t_cov) @ torch.inverse(mat_sqrt_quirk(s_cov))
    out = (m @ s_n) * t_std + t_mean
    return out.view(source.shape)


def random_noise(x: Tensor, noise: Tensor) -> Tensor:
    """modules.py:7-12 with the N(0, std) draw passed in."""
    return torch.clip(x + noise, -1.0, 1.0)


def gaussian_kernel1d(ksize: int = 5, sigma: Optional[float] = None, dtype=torch.float32) -> Tensor:
    """torchvision gaussian_blur default sigma = 0.3*((k-1)*0.5-1)+0.8 (=1.1 for k=5)."""
    if sigma is None:
        sigma = 0.3 * ((ksize - 1) * 0.5 - 1) + 0.8
    half = (ksize - 1) * 0.5
    xs = torch.linspace(-half, half, steps=ksize, dtype=dtype)
    pdf = torch.exp(-0.5 * (xs / sigma) ** 2)
    return pdf / pdf.sum()


def random_blur(x: Tensor, ksize: int = 5) -> Tensor:
    """modules.py:15-16: 5x5 Gaussian (sigma 1.1), reflect padding, depthwise."""
    k1 = gaussian_kernel1d(ksize, dtype=x.dtype)
    k2 = torch.outer(k1, k1)
    C = x.shape[1]
    p = ksize // 2
    xp = F.pad(x, (p, p, p, p), mode="reflect")
    return F.conv2d(xp, k2.expand(C, 1, ksize, ksize).contiguous(), groups=C)


def _nearest_idx(in_size: int, out_size: int, scale: Optional[float]) -> Tensor:
    """ATen nearest_neighbor_compute_source_index: floor(dst * scale) clamped, scale in fp32
    (= 1/scale_factor when a scale_factor was given, else in/out)."""
    s = torch.tensor(scale if scale is not None else in_size / out_size, dtype=torch.float32)
    idx = torch.floor(torch.arange(out_size, dtype=torch.float32) * s).to(torch.int64)
    return idx.clamp(max=in_size - 1)


def downscale(x: Tensor, bottleneck_scale: float = 0.75) -> Tensor:
    """modules.py:19-21: nearest down by 0.75 (scale_factor form) then nearest back (size form)."""
    H, W = x.shape[-2:]
    h, w = int(math.floor(H * bottleneck_scale)), int(math.floor(W * bottleneck_scale))
    iy = _nearest_idx(H, h, 1.0 / bottleneck_scale)
    ix = _nearest_idx(W, w, 1.0 / bottleneck_scale)
    down = x.index_select(-2, iy).index_select(-1, ix)
    jy = _nearest_idx(h, H, None)
    jx = _nearest_idx(w, W, None)
    return down.index_select(-2, jy).index_select(-1, jx)


# ----------------------------------------------------------------------------------------
# a17 (secondary): SFConv frequency branch      (model/efficientnet/exp.py:55-65; model/resnet/exp.py:44-54)
# ----------------------------------------------------------------------------------------
def sfconv_freq_branch(x: Tensor, freq_w: Tensor, out_size: Optional[Sequence[int]] = None,
                       norm: Optional[str] = "ortho") -> Tensor:
    size = x.shape[-2:]
    f = F.conv2d(cat_rfft2(x, norm), freq_w)
    y = irfft2_from_cat(f, size, norm)
    if out_size is not None and tuple(out_size) != tuple(y.shape[-2:]):
        y = F.adaptive_avg_pool2d(y, tuple(out_size))
    return y


# ----------------------------------------------------------------------------------------
# isolated recon path: one fwd(+bwd) pass on cached backbone features
# ----------------------------------------------------------------------------------------
def recon_path_forward(arch: str, x: Tensor, feat: Tensor, emb: Tensor, labels: Tensor,
                       params: Dict[str, Tensor], lambdas: Optional[Dict[str, float]] = None):
    """decoder -> attention -> rec tail -> triplet, combined with the engine's pass-1 weights
    (engine/abstract_engine.py:233-267, real rows only for rec/freq).  Dropout disabled.
    Returns (loss, dict of intermediates)."""
    lam = {"triplet": 0.1, "recons": 0.1, "freq": 1.0, "mask": 0.1}
    if lambdas:
        lam.update(lambdas)
    act = DECODER_ACT[arch]
    dec_outs = decoder(feat, params, arch)
    dec_last = dec_outs[-1]
    att_out, fmask, smask = attention(dec_last.detach(), x, emb, params, act, training=True)
    rec, spatial, freq = recon_tail(dec_last, x)
    n_real = int((labels == 0).sum())
    tri_feats = [feat.mean(dim=(-2, -1))] + [d.mean(dim=(-2, -1)) for d in dec_outs[:TRIPLET_DEC[arch]]]
    tri = sum(aw_triplet_loss(f_, labels) for f_ in tri_feats)
    loss = (lam["mask"] * fmask.mean() + lam["mask"] * smask.mean() + lam["triplet"] * tri
            + lam["recons"] * spatial[:n_real].mean() + lam["freq"] * freq[:n_real].mean())
    return loss, {"rec": rec, "spatial": spatial, "freq": freq, "freq_mask": fmask, "spat_mask": smask,
                  "att_out": att_out, "dec_outs": dec_outs, "triplet": tri}
