"""Run-to-run determinism of every op of the UDR18 model (full-model fixture shapes) on the GPU.

    UD_RUNS=100 python tools/determinism_check.py

Each op (whole blocks, their leaf modules, and every hot-path kernel on captured inputs) runs UD_RUNS times back to back
on identical inputs WITHOUT host synchronisation in between; the (sum, |.|-sum) checksums of all outputs and gradients are
compared at the end.  Written to find why two scalar gradients of the toy-size model test moved by 4 % in ~1 run of 8:
every kernel of this repo is bit-identical run to run; cuDNN's ConvTranspose2d forward and convolution backward are not
(atomics), and that jitter flips a channel-argmax of the dynamic filters (profiles/r02_determinism_udr18.txt)."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch, torch.nn.functional as F
import procedural as P

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
F.dropout = lambda t, p=0.5, training=True, inplace=False: t * 1.0
from unidefense_b200 import ops
from unidefense_b200.model import load_model
fix = torch.load(os.path.join(ROOT, "tests/golden/full_r18.pt"), weights_only=False)
torch.manual_seed(20260117)
model = load_model("UDR18")(num_classes=2, drop_rate=0.0)
P.fill_state_dict_(model, prefix_filter=None, salt=5)
model = model.cuda().train()
cap = {}
orig = model.attention
def spy(pred, x, embedding):
    cap["args"] = (pred.detach().clone(), x.detach().clone(), embedding.detach().clone())
    return orig(pred, x, embedding)
model.attention = spy
x = P.tensor_for("in:full_x_r18", (fix["N"], 3, fix["R"], fix["R"]), "unit").cuda()
model(x)
model.attention = orig
pred, xin, emb0 = cap["args"]
print("attention inputs:", tuple(pred.shape), tuple(xin.shape), tuple(emb0.shape))
g = torch.Generator(device="cuda").manual_seed(1)
size = emb0.shape[-2:]
N, C = emb0.shape[:2]
go = torch.randn(emb0.shape, device="cuda", generator=g)
gfm = torch.randn(N, 1, size[0], size[1] // 2 + 1, device="cuda", generator=g)
gsm = torch.randn(N, 1, size[0], size[1], device="cuda", generator=g)
params = {n: p for n, p in model.named_parameters() if n.split(".")[0] in ("freq_filter", "spat_filter", "fuse_coef")}
R = int(os.environ.get("UD_RUNS", "60"))


def tally(name, fn):
    """fn() -> dict of tensors; counts distinct outcomes per tensor over R runs."""
    seen = collections.defaultdict(dict)
    rec = []
    for _ in range(R):                      # NO host synchronisation inside the loop: the GPU runs the ops back to back
        for p in params.values():
            p.grad = None
        out = fn()
        rec.append({k: torch.stack([v.double().sum(), v.double().abs().sum()]) for k, v in out.items() if v is not None})
    torch.cuda.synchronize()
    for r in rec:
        for k, v in r.items():
            key = tuple(v.tolist())
            seen[k][key] = seen[k].get(key, 0) + 1
    for k, d in seen.items():
        if len(d) > 1:
            items = sorted(d.items(), key=lambda t: -t[1])
            spread = max(abs(a[0][1] - items[0][0][1]) for a in items) / max(items[0][0][1], 1e-30)
            print(f"  [{name}] {k}: {len(d)} distinct outcomes, counts {[c for _, c in items][:6]}, rel spread of |.|-sum {spread:.2e}")
    print(f"[{name}] done ({sum(len(d) > 1 for d in seen.values())} of {len(seen)} tensors vary)")


def whole():
    e = emb0.clone().requires_grad_()
    r = model.attention(pred, xin, e)
    (r["out"] * go).sum().add((r["freq_mask"] * gfm).sum()).add((r["spat_mask"] * gsm).sum()).backward()
    d = {"out": r["out"].detach(), "fm": r["freq_mask"].detach(), "sm": r["spat_mask"].detach(), "g_emb": e.grad}
    d.update({"g:" + n: p.grad for n, p in params.items()})
    return d


sd, fd = ops.attn_prep(pred.float(), xin.float(), size, model.freq_norm)


def freq_path():
    e = emb0.clone().requires_grad_()
    ef = ops.rfft2_cat(e, model.freq_norm)
    fo = model.freq_filter(ef, fd)
    ff = ops.irfft2_cat(fo["out"], size, model.freq_norm)
    (ff * go).sum().add((fo["mask"] * gfm).sum()).backward()
    d = {"ef": ef.detach(), "ff": ff.detach(), "fm": fo["mask"].detach(), "g_emb": e.grad}
    d.update({"g:" + n: p.grad for n, p in params.items() if n.startswith("freq_filter")})
    return d


def fft_only():
    e = emb0.clone().requires_grad_()
    ef = ops.rfft2_cat(e, model.freq_norm)
    ff = ops.irfft2_cat(ef * 1.5, size, model.freq_norm)
    (ff * go).sum().backward()
    return {"ef": ef.detach(), "ff": ff.detach(), "g_emb": e.grad}


def freq_filter_only():
    ef = ops.rfft2_cat(emb0, model.freq_norm).detach().requires_grad_()
    fo = model.freq_filter(ef, fd)
    (fo["out"] * torch.ones_like(fo["out"])).sum().add((fo["mask"] * gfm).sum()).backward()
    d = {"out": fo["out"].detach(), "fm": fo["mask"].detach(), "g_ef": ef.grad}
    d.update({"g:" + n: p.grad for n, p in params.items() if n.startswith("freq_filter")})
    return d


def spat_path():
    e = emb0.clone().requires_grad_()
    sm = model.spat_filter.mask_only(e, sd)
    (sm * gsm).sum().backward()
    d = {"sm": sm.detach(), "g_emb": e.grad}
    d.update({"g:" + n: p.grad for n, p in params.items() if n.startswith("spat_filter")})
    return d


def fuse_only():
    e = emb0.clone().requires_grad_()
    sm = torch.sigmoid(gsm).requires_grad_()
    ff = (emb0 * 0.5 + 0.1).requires_grad_()
    o = ops.attn_fuse(e, sm, ff, None, model.fuse_coef)
    (o * go).sum().backward()
    return {"o": o.detach(), "g_emb": e.grad, "g_sm": sm.grad, "g_ff": ff.grad, "g:fuse_coef": model.fuse_coef.grad}


def proj_only(kind):
    conv = (model.spat_filter if kind == "spat" else model.freq_filter).layer1[0]
    xin_ = emb0 if kind == "spat" else ops.rfft2_cat(emb0, model.freq_norm).detach()
    def f():
        xx = xin_.clone().requires_grad_()
        y, pm, p2 = ops.proj_conv(xx, conv.weight, want_stats=True)
        (y * torch.ones_like(y)).sum().backward()
        return {"y": y.detach(), "pm": pm, "p2": p2, "gx": xx.grad, "gw": conv.weight.grad}
    return f


labels = fix["labels"].cuda()


def attention_and_head():
    e = emb0.clone().requires_grad_()
    r = model.attention(pred, xin, e)
    feat = F.adaptive_avg_pool2d(r["out"], 1).flatten(1)
    feat = model.bottleneck(feat)
    logits = model.classifier(feat)
    loss = F.cross_entropy(logits, labels) + 0.1 * r["freq_mask"].mean() + 0.1 * r["spat_mask"].mean()
    loss.backward()
    d = {"logits": logits.detach(), "g_emb": e.grad}
    d.update({"g:" + n: p.grad for n, p in params.items()})
    return d


# ---- decoder side: whole blocks, their leaf convs, in_act / tanh on captured inputs; extractor; recon tail ----
cap2 = {}
mods = dict(model.named_modules())
hooks = [mods[n].register_forward_hook(lambda mod, inp, out, n=n: cap2.__setitem__(n, (inp[0].detach().clone(), out)))
         for n in ("dec_block1", "dec_block2", "extractor")]
for n, m in model.named_modules():
    if n.split(".")[0] in ("dec_block1", "dec_block2") and isinstance(m, (torch.nn.Conv2d, torch.nn.ConvTranspose2d)):
        hooks.append(m.register_forward_hook(lambda mod, inp, out, n=n: cap2.__setitem__(n, (inp[0].detach().clone(), None))))
model(x)
for h in hooks:
    h.remove()


def mod_fn(m, xin_, with_mean=False):
    def f():
        xx = xin_.clone().requires_grad_()
        y = m(xx)
        if isinstance(y, (tuple, list)):
            y = y[-1]
        for p in m.parameters():
            p.grad = None
        y.backward(torch.ones_like(y) * 0.37)
        d = {"y": y.detach(), "gx": xx.grad}
        d.update({"g:" + k: p.grad for k, p in m.named_parameters() if p.grad is not None})
        return d
    return f


for n, (xin_, _) in cap2.items():
    try:
        tally(n + " " + type(mods[n]).__name__, mod_fn(mods[n], xin_))
    except Exception as ex:      # noqa: BLE001
        print(f"[{n}] raised {type(ex).__name__}: {ex}")
for n in ("dec_block1", "dec_block2"):
    blk = mods[n]
    kids = list(blk)
    for i, m in enumerate(kids):
        if isinstance(m, torch.nn.InstanceNorm2d):
            xin_ = cap2[f"{n}.{i - 1}"][0]
            conv = kids[i - 1]
            with torch.no_grad():
                pre = conv(xin_).float()
            act = "relu"
            def f(pre=pre, m=m):
                xx = pre.clone().requires_grad_()
                m.weight.grad = None; m.bias.grad = None
                y, ym = ops.in_act(xx, m.weight, m.bias, "relu", m.eps, want_mean=True)
                (y * 0.37).sum().add(ym.sum()).backward()
                return {"y": y.detach(), "ym": ym.detach(), "gx": xx.grad, "gw": m.weight.grad, "gb": m.bias.grad}
            tally(f"in_act after {n}.{i - 1} {tuple(pre.shape)}", f)
dec2 = cap2["dec_block2"][1].detach()


def tail():
    d = dec2.clone().requires_grad_()
    rec, sp, fr = ops.recon_tail(d, x.float(), model.freq_norm)
    (0.1 * sp[:2].mean() + fr[:2].mean()).backward()
    return {"rec": rec.detach(), "sp": sp.detach(), "fr": fr.detach(), "gd": d.grad}


def prep():
    sd_, fd_ = ops.attn_prep(dec2, x.float(), size, model.freq_norm)
    return {"sd": sd_, "fd": fd_}


tally("recon_tail", tail)
tally("attn_prep", prep)

for name, fn in [("attention + head", attention_and_head), ("whole attention", whole), ("freq path", freq_path), ("fft only", fft_only), ("freq filter only", freq_filter_only),
                 ("spat path", spat_path), ("fuse only", fuse_only), ("proj spat 3x3", proj_only("spat")),
                 ("proj freq 1x1", proj_only("freq"))]:
    try:
        tally(name, fn)
    except Exception as ex:      # noqa: BLE001
        print(f"[{name}] raised {type(ex).__name__}: {ex}")
