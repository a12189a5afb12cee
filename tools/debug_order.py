"""temporary: find what state test_dyfi_gpu leaves behind that changes the UDR18 gradients."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch, torch.nn.functional as F
import procedural as P
import pytest

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
F.dropout = lambda t, p=0.5, training=True, inplace=False: t * 1.0


def run():
    from unidefense_b200 import ops
    from unidefense_b200.model import load_model
    fix = torch.load(os.path.join(ROOT, "tests/golden/full_r18.pt"), weights_only=False)
    torch.manual_seed(20260117)
    model = load_model("UDR18")(num_classes=2, drop_rate=0.0)
    P.fill_state_dict_(model, prefix_filter=None, salt=5)
    model = model.cuda().train()
    acts = {}
    def hook(name):
        def f(mod, inp, out):
            def grab(o, k):
                if torch.is_tensor(o): acts[k] = o.detach().clone()
                elif isinstance(o, dict):
                    for kk, v in o.items(): grab(v, f"{k}.{kk}")
                elif isinstance(o, (list, tuple)):
                    for i, v in enumerate(o): grab(v, f"{k}.{i}")
            grab(out, name)
        return f
    for n, m in model.named_modules():
        if n and n.count(".") <= 1:
            m.register_forward_hook(hook(n))
    x = P.tensor_for("in:full_x_r18", (fix["N"], 3, fix["R"], fix["R"]), "unit").cuda()
    labels = fix["labels"].cuda()
    out = model(x)
    ld = out["loss_dict"]
    nr = fix["N"] // 2
    tri = sum(ops.triplet_loss(f, labels) for f in ld["triplet"])
    loss = (F.cross_entropy(out["cls_out"], labels) + 0.1 * ld["freq_mask"].mean() + 0.1 * ld["spat_mask"].mean()
            + 0.1 * tri + 0.1 * ld["spatial"][:nr].mean() + 1.0 * ld["freq"][:nr].mean())
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    return acts, grads, fix


a1, g1, fix = run()
a1b, g1b, _ = run()
print("run-to-run (fresh, fresh):", max(float((g1[k] - g1b[k]).abs().max()) for k in g1))
which = sys.argv[1] if len(sys.argv) > 1 else "tests/test_dyfi_gpu.py"
pytest.main([which, "-q", "-m", "gpu", "-x", "-p", "no:cacheprovider"] + sys.argv[2:])
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
F.dropout = lambda t, p=0.5, training=True, inplace=False: t * 1.0
a2, g2, _ = run()
print("=== forward activations that differ after", which)
for k in a1:
    if k in a2 and a1[k].shape == a2[k].shape and a1[k].dtype.is_floating_point:
        d = float((a1[k] - a2[k]).abs().max()); s = float(a1[k].abs().max())
        if d > 1e-6 * max(s, 1e-6):
            print(f"  {k}: max abs diff {d:.3e} (scale {s:.3e})")
print("=== parameter gradients that differ")
for k in g1:
    d = float((g1[k] - g2[k]).abs().max()); s = float(g1[k].abs().max())
    if d > 1e-5 * max(s, 1e-12):
        print(f"  {k}: max abs diff {d:.3e} (scale {s:.3e}) norm {float(g1[k].norm()):.5f} -> {float(g2[k].norm()):.5f}  ref {fix['param_grads'][k]['norm'] if fix['param_grads'].get(k) else None}")
print("flags now: cudnn.allow_tf32", torch.backends.cudnn.allow_tf32, "matmul", torch.backends.cuda.matmul.allow_tf32,
      "fp32 matmul precision", torch.get_float32_matmul_precision(), "cudnn.benchmark", torch.backends.cudnn.benchmark,
      "deterministic", torch.backends.cudnn.deterministic)
