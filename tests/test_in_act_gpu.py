"""a2 parity: InstanceNorm+activation and tanh epilogues vs oracle / reference fixtures."""
import pytest
import torch

from oracle import recon_path as O

pytestmark = pytest.mark.gpu

# (N, C, H, W): register path (cs=1), cluster paths (cs=2,4,8), generic (odd E), tiny
SHAPES = [(2, 4, 6, 6), (2, 5, 24, 24), (3, 8, 48, 48), (2, 6, 96, 96), (2, 5, 192, 192), (1, 3, 190, 190),
          (2, 3, 95, 95), (2, 7, 5, 3), (1, 2, 128, 128), (2, 3, 64, 64), (1, 1, 300, 300)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("act", ["swish", "relu"])
def test_in_act_fwd_bwd(shape, act):
    from unidefense_b200 import ops
    N, C, H, W = shape
    g = torch.Generator().manual_seed(N * 1000 + C * 10 + H)
    x = torch.randn(shape, generator=g) * 1.5 + 0.3
    gamma = torch.rand(C, generator=g) + 0.5
    beta = torch.randn(C, generator=g) * 0.1
    gy = torch.randn(shape, generator=g)
    gm = torch.randn(N, C, generator=g)
    xc, gc, bc = x.cuda().requires_grad_(), gamma.cuda().requires_grad_(), beta.cuda().requires_grad_()
    y, ym = ops.in_act(xc, gc, bc, act, want_mean=True)
    ((y * gy.cuda()).sum() + (ym * gm.cuda()).sum()).backward()

    x64, g64, b64 = x.double().requires_grad_(), gamma.double().requires_grad_(), beta.double().requires_grad_()
    y64 = O.instance_norm_act(x64, g64, b64, act)
    ((y64 * gy.double()).sum() + (y64.mean(dim=(-2, -1)) * gm.double()).sum()).backward()

    def close(a, b, rtol=1e-4):
        torch.testing.assert_close(a.detach().cpu().double(), b.detach(), rtol=rtol, atol=rtol * float(b.abs().max()) * 0.1 + 1e-9)

    close(y, y64)
    close(ym, y64.mean(dim=(-2, -1)))
    # relu kink: elements with |z| at fp32 noise level may flip; none expected for random data
    close(xc.grad, x64.grad, rtol=2e-4)
    # parameter grads are sums of N*H*W signed terms: tolerance relative to the term scale
    n_terms = N * H * W
    for got, want in ((gc.grad, g64.grad), (bc.grad, b64.grad)):
        torch.testing.assert_close(got.detach().cpu().double(), want, rtol=2e-4, atol=1e-6 * n_terms ** 0.5 * 3 + 1e-6)


def test_in_act_reference_fixture(golden_ops):
    from unidefense_b200 import ops
    for c in golden_ops["in_act"]:
        x, g, b = c["x"].cuda().requires_grad_(), c["gamma"].cuda().requires_grad_(), c["beta"].cuda().requires_grad_()
        y = ops.in_act(x, g, b, c["act"])
        torch.testing.assert_close(y.detach().cpu(), c["y"], rtol=1e-4, atol=1e-5)
        (y * c["gy"].cuda()).sum().backward()
        torch.testing.assert_close(x.grad.cpu(), c["gx"], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(g.grad.cpu(), c["ggamma"], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(b.grad.cpu(), c["gbeta"], rtol=1e-4, atol=1e-5)


def test_in_act_no_affine_and_empty():
    from unidefense_b200 import ops
    x = torch.randn(2, 3, 8, 8, device="cuda", requires_grad=True)
    y = ops.in_act(x, None, None, "relu")
    ref = torch.relu(torch.nn.functional.instance_norm(x.detach()))
    torch.testing.assert_close(y, ref, rtol=1e-4, atol=1e-5)
    y.sum().backward()
    assert x.grad.shape == x.shape
    e = ops.in_act(torch.zeros(0, 3, 8, 8, device="cuda"), None, None, "swish")
    assert e.shape == (0, 3, 8, 8)


def test_tanh():
    from unidefense_b200 import ops
    x = torch.randn(3, 3, 37, 41, device="cuda", requires_grad=True)
    y = ops.tanh(x)
    torch.testing.assert_close(y, torch.tanh(x.detach()), rtol=1e-5, atol=1e-6)
    gy = torch.randn_like(y)
    (y * gy).sum().backward()
    torch.testing.assert_close(x.grad, gy * (1 - torch.tanh(x.detach()) ** 2), rtol=1e-4, atol=1e-6)


# bf16 I/O variants: shapes covering the register path (small / big slots), clusters (cs = 2..8), half-filled last slots,
# the generic kernel (HW not a multiple of 8) and the shipped decoder planes (20 x 192^2, 40 x 96^2, 80 x 48^2, 80 x 24^2)
BF16_SHAPES = [(2, 4, 8, 8), (2, 5, 24, 24), (3, 8, 48, 48), (2, 6, 96, 96), (2, 20, 192, 192), (1, 3, 190, 190),
               (2, 3, 95, 95), (2, 7, 5, 3), (1, 2, 128, 128), (2, 3, 64, 64), (1, 1, 300, 300), (2, 40, 96, 96)]


@pytest.mark.parametrize("shape", BF16_SHAPES)
@pytest.mark.parametrize("act", ["swish", "relu"])
def test_in_act_bf16_io_equals_cast_fp32_cast(shape, act):
    """bf16 in / bf16 out == (bf16 -> fp32) -> fp32 kernel -> (fp32 -> bf16).  The two kernels split a plane into
    different slabs (8- vs 4-element vectors), so their fp32 statistics differ in the last bit and a result that sits
    on a bf16 rounding boundary may land on the neighbouring bf16 value: equal up to ONE bf16 ulp, on < 1 % of the
    elements."""
    from unidefense_b200 import ops
    N, C, H, W = shape
    g = torch.Generator().manual_seed(N * 1000 + C * 10 + H + 1)
    x = (torch.randn(shape, generator=g) * 1.5 + 0.3).to(torch.bfloat16)
    gamma = torch.rand(C, generator=g) + 0.5
    beta = torch.randn(C, generator=g) * 0.1
    gy = torch.randn(shape, generator=g).to(torch.bfloat16)
    gm = torch.randn(N, C, generator=g)
    xb, gb, bb = x.cuda().requires_grad_(), gamma.cuda().requires_grad_(), beta.cuda().requires_grad_()
    y, ym = ops.in_act(xb, gb, bb, act, want_mean=True)
    assert y.dtype == torch.bfloat16 and ym.dtype == torch.float32
    ((y.float() * gy.cuda().float()).sum() + (ym * gm.cuda()).sum()).backward()
    assert xb.grad.dtype == torch.bfloat16

    xf, gf, bf = x.cuda().float().requires_grad_(), gamma.cuda().requires_grad_(), beta.cuda().requires_grad_()
    yf, ymf = ops.in_act(xf, gf, bf, act, want_mean=True)
    ((yf * gy.cuda().float()).sum() + (ymf * gm.cuda()).sum()).backward()
    def one_ulp(a, b, what):
        a, b = a.detach().float(), b.detach().to(torch.bfloat16).float()
        diff = (a - b).abs()
        # (gx = k (gz - S1/E - xh S2/E) cancels: near its zeros the 1e-7 difference of the sums is many ulps of the result)
        floor = 1e-5 * float(b.abs().max())
        assert bool((diff <= 2.0 ** -7 * torch.maximum(a.abs(), b.abs()) + floor).all()), f"{what}: more than one bf16 ulp apart"
        assert float((diff > 0).float().mean()) < 0.01, f"{what}: {float((diff > 0).float().mean()):.4f} of the elements differ"

    one_ulp(y, yf, "y")
    torch.testing.assert_close(ym, ymf, rtol=1e-5, atol=1e-6)
    one_ulp(xb.grad, xf.grad, "gx")
    n_terms = N * H * W
    torch.testing.assert_close(gb.grad, gf.grad, rtol=1e-4, atol=1e-6 * n_terms ** 0.5 * 3 + 1e-6)
    torch.testing.assert_close(bb.grad, bf.grad, rtol=1e-4, atol=1e-6 * n_terms ** 0.5 * 3 + 1e-6)


def test_decoder_block_keeps_bf16_under_autocast():
    """Under bf16 autocast the (conv, InstanceNorm, act) chain stays bf16 end to end: no fp32 tensor between the
    convolutions, and the result equals the fp32-epilogue composition it replaces."""
    from unidefense_b200.model.modules import make_decoder_block
    import torch.nn as nn
    from unidefense_b200.model.modules import MemoryEfficientSwish
    torch.manual_seed(0)
    blk = make_decoder_block([("c", 16, 8), ("t", 8, 8), ("c", 8, 8)], nn.InstanceNorm2d, MemoryEfficientSwish, True, False).cuda()
    x = torch.randn(2, 16, 24, 24, device="cuda")
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y, ym = blk.forward_with_mean(x)
    assert y.dtype == torch.bfloat16 and ym.dtype == torch.float32 and y.shape == (2, 8, 48, 48)
    from unidefense_b200 import ops
    with torch.autocast("cuda", dtype=torch.bfloat16):
        z = x
        mods = list(blk)
        for i in range(0, len(mods), 3):
            z = ops.in_act(mods[i](z).float(), mods[i + 1].weight, mods[i + 1].bias, "swish", mods[i + 1].eps)
    d = (y.float() - z.to(torch.bfloat16).float()).abs()            # (three chained stages: a 1-ulp flip propagates)
    assert float(d.max()) <= 0.05 * float(z.abs().max()) and float((d > 0).float().mean()) < 0.2
