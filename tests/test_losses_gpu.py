"""a9-a11 parity: fused triplet / factorization / mask-KL losses (value + gradient)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import recon_path as O

pytestmark = pytest.mark.gpu


def close(a, b, rtol=1e-4, atol=None):
    b = b.detach()
    scale = float(b.abs().max()) if b.numel() else 1.0
    atol = (1e-5 * max(scale, 1e-6)) if atol is None else atol
    torch.testing.assert_close(a.detach().cpu().to(b.dtype), b, rtol=rtol, atol=atol)


def test_triplet_reference_fixture(golden_ops):
    from unidefense_b200 import ops
    for c in golden_ops["triplet"]:
        f = c["feat"].cuda().requires_grad_()
        l = ops.triplet_loss(f, c["labels"].cuda())
        close(l, c["loss"])
        (3.0 * l).backward()
        dup = bool((torch.cdist(c["feat"], c["feat"]) + torch.eye(len(c["feat"])) < 1e-6).any())
        if not dup:   # clamp(1e-12).sqrt() at an exact duplicate: fp32 rounding decides the branch
            close(f.grad, 3.0 * c["gfeat"], rtol=1e-3, atol=1e-5 * float(c["gfeat"].abs().max()) * 3)


@pytest.mark.parametrize("N,nr,c", [(4, 2, 16), (32, 16, 160), (32, 16, 80), (32, 16, 40), (64, 32, 1024), (20, 10, 448),
                                    (6, 1, 8), (5, 5, 8), (128, 64, 32), (3, 2, 1)])
def test_triplet_vs_oracle(N, nr, c):
    from unidefense_b200 import ops
    g = torch.Generator().manual_seed(N * 7 + c)
    feat = torch.randn(N, c, generator=g) * (1.0 / c ** 0.5)
    labels = torch.tensor([0] * nr + [1] * (N - nr))
    f = feat.cuda().requires_grad_()
    l = ops.triplet_loss(f, labels.cuda())
    l.backward()
    f64 = feat.double().requires_grad_()
    l64 = O.aw_triplet_loss(f64, labels)
    l64.backward()
    close(l, l64, rtol=1e-4)
    close(f.grad, f64.grad, rtol=2e-3, atol=2e-4 * float(f64.grad.abs().max()) + 1e-9)


def test_factorization_reference_fixture(golden_ops):
    from unidefense_b200 import ops
    for c in golden_ops["factorization"]:
        a = c["a"].cuda().requires_grad_()
        l = ops.factorization_loss(a, c["b"].cuda())
        close(l, c["loss"])
        l.backward()
        # N=2 makes the z-scores constant (+-1/sqrt2): the true gradient is ~0 and both sides hold fp32
        # cancellation residue of O(eps * |dL/dz|) -> absolute floor
        close(a.grad, c["ga"], rtol=1e-3, atol=1e-4 * float(c["ga"].abs().max()) + 5e-7)


@pytest.mark.parametrize("N,F", [(32, 1792), (64, 2048), (4, 512), (2, 5), (20, 100), (128, 70)])
def test_factorization_vs_oracle(N, F):
    from unidefense_b200 import ops
    g = torch.Generator().manual_seed(N + F)
    a = torch.randn(N, F, generator=g) * 0.7 + 0.1
    b = a * 0.5 + torch.randn(N, F, generator=g) * 0.5
    ac = a.cuda().requires_grad_()
    l = ops.factorization_loss(ac, b.cuda())
    (2.0 * l).backward()
    a64 = a.double().requires_grad_()
    l64 = O.factorization_loss(a64, b.double())
    (2.0 * l64).backward()
    close(l, l64, rtol=1e-4)
    close(ac.grad, a64.grad, rtol=2e-3, atol=2e-4 * float(a64.grad.abs().max()) + 5e-7)


def test_factorization_errors():
    from unidefense_b200 import ops
    with pytest.raises(RuntimeError):
        ops.factorization_loss(torch.zeros(1, 8, device="cuda"), torch.zeros(1, 8, device="cuda"))   # N >= 2


def test_mask_kl_reference_fixture(golden_ops):
    from unidefense_b200 import ops
    for c in golden_ops["mask_kl"]:
        p = c["pred"].cuda().requires_grad_()
        l = ops.mask_kl_loss(p, c["gt"].cuda())
        close(l, c["loss"], rtol=1e-3, atol=1e-8)
        l.backward()
        close(p.grad, c["gpred"], rtol=1e-3, atol=1e-8)


@pytest.mark.parametrize("shape", [(32, 1, 12, 7), (32, 1, 12, 12), (64, 1, 8, 5), (4, 1, 24, 24), (1, 1, 1, 1), (3, 1, 1, 300)])
def test_mask_kl_vs_oracle(shape):
    from unidefense_b200 import ops
    g = torch.Generator().manual_seed(sum(shape))
    p = torch.sigmoid(torch.randn(shape, generator=g))
    q = torch.sigmoid(torch.randn(shape, generator=g))
    pc = p.cuda().requires_grad_()
    l = ops.mask_kl_loss(pc, q.cuda())
    l.backward()
    p64 = p.double().requires_grad_()
    l64 = O.mask_kl_loss(p64, q.double())
    l64.backward()
    close(l, l64, rtol=1e-3, atol=1e-7)
    close(pc.grad, p64.grad, rtol=1e-3, atol=1e-7)


def test_cross_entropy_and_bce_match_torch():
    """a12: the engine's two classification-loss branches (engine/abstract_engine.py:256-259)."""
    from unidefense_b200.loss import get_loss
    g = torch.Generator().manual_seed(3)
    for N, K in [(4, 2), (32, 2), (7, 5), (64, 2)]:
        z = (torch.randn(N, K, generator=g) * 3).cuda().requires_grad_()
        t = torch.randint(0, K, (N,), generator=g).cuda()
        loss = get_loss("cross_entropy", "cuda")(z, t)
        (gz,) = torch.autograd.grad(loss * 1.7, [z])
        z64 = z.detach().double().cpu().requires_grad_()
        ref = F.cross_entropy(z64, t.cpu())
        (gr,) = torch.autograd.grad(ref * 1.7, [z64])
        torch.testing.assert_close(loss.double().cpu(), ref.detach(), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(gz.double().cpu(), gr, rtol=1e-4, atol=1e-7)
        zb = (torch.randn(N, generator=g) * 4).cuda().requires_grad_()
        tb = torch.randint(0, 2, (N,), generator=g).float().cuda()
        lb = get_loss("bce", "cuda")(zb, tb)
        (gb,) = torch.autograd.grad(lb, [zb])
        zb64 = zb.detach().double().cpu().requires_grad_()
        rb = F.binary_cross_entropy_with_logits(zb64, tb.double().cpu())
        (grb,) = torch.autograd.grad(rb, [zb64])
        torch.testing.assert_close(lb.double().cpu(), rb.detach(), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(gb.double().cpu(), grb, rtol=1e-4, atol=1e-7)
