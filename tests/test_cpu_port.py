"""The CPU arm of bench.py (oracle/ref_model.py: the drop-in's stock-torch backbone + the oracle's hot path) computes
what the reference computes: checked against the whole-model fixtures recorded from the reference classes
(tests/golden/full_*.pt).  This is what makes `cpu_baseline.kind == "port"` a faithful stand-in on the GPU box,
where /root/reference does not exist.  Also pins the DFT-by-GEMM matrices of the SFConv fast path."""
import os

import pytest
import torch
import torch.nn.functional as F

import procedural as P
from oracle import recon_path as O
from oracle import ref_model

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("arch", ["r18", "r50", "eb4"])
def test_cpu_port_matches_reference_fixture(arch, monkeypatch):
    fix = torch.load(os.path.join(GOLDEN, f"full_{arch}.pt"), weights_only=False)
    kw = dict(num_classes=2, drop_rate=0.0)
    if arch == "eb4":
        kw["drop_connect_rate"] = 0.0
    if arch == "r50":
        kw["extractor"] = "resnet50"
    model = ref_model.build(arch, **kw)
    P.fill_state_dict_(model, salt=5)
    model.train()
    monkeypatch.setattr(F, "dropout", lambda t, p=0.5, training=True, inplace=False: t * 1.0)
    x = P.tensor_for(f"in:full_x_{arch}", (fix["N"], 3, fix["R"], fix["R"]), "unit")
    out = model(x)
    ld = out["loss_dict"]
    noise = fix["noise"]

    def near(a, b, nz, what):
        tol = 1e-4 * max(float(b.abs().max()), 1e-6) + 8.0 * nz
        err = float((a.detach() - b).abs().max())
        assert err <= tol, f"{what}: max abs err {err:.3e} > {tol:.3e}"

    near(ld["spatial"], fix["spatial"], noise["spatial"], "spatial")
    near(ld["freq"], fix["freq"], noise["freq"], "freq")
    near(ld["freq_mask"], fix["freq_mask"], noise["freq_mask"], "freq_mask")
    near(ld["spat_mask"], fix["spat_mask"], noise["spat_mask"], "spat_mask")
    near(out["rec"][:, :, ::7, ::5], fix["rec_sample"], noise["rec_sample"], "rec")
    near(ld["factorization"], fix["factorization"], noise["factorization"], "factorization")
    near(out["cls_out"], fix["cls_out"], noise["cls_out"], "cls_out")
    lam = dict(mask=0.1, triplet=0.1, recons=0.1, freq=1.0)
    loss = ref_model.pass1_loss(out, fix["labels"], fix["N"] // 2, lam)
    near(loss, fix["loss"], noise["loss"], "loss")


def test_dft_gemm_matrices_match_torch_fft():
    from unidefense_b200.model import sfconv
    for (h, w) in [(12, 12), (24, 24), (48, 48), (9, 7), (8, 6), (5, 8), (64, 64), (1, 1)]:
        for norm in ("ortho", None):
            N, C = 2, 4
            wh = w // 2 + 1
            L, R, Li, A = [t.double() for t in sfconv._dft_mats(h, w, norm, "cpu")]
            g = torch.Generator().manual_seed(h * 100 + w)
            x = torch.randn(N, C, h, w, generator=g, dtype=torch.float64).contiguous(memory_format=torch.channels_last)
            V = torch.matmul(L, x.permute(0, 2, 3, 1).reshape(N, h, w * C))
            planar = torch.matmul(R, V.reshape(N * h, 2 * w, C)).reshape(N, h, wh, 2 * C).permute(0, 3, 1, 2)
            torch.testing.assert_close(planar, O.cat_rfft2(x, norm), rtol=1e-5, atol=2e-6 * max(h * w, 1) ** 0.5)
            Q = torch.randn(N, 2 * C, h, wh, generator=g, dtype=torch.float64).contiguous(memory_format=torch.channels_last)
            G = torch.matmul(Li, Q.permute(0, 2, 3, 1).reshape(N, h, wh * 2 * C))
            y = torch.matmul(A, G.reshape(N * h, 4 * wh, C)).reshape(N, h, w, C).permute(0, 3, 1, 2)
            torch.testing.assert_close(y, O.irfft2_from_cat(Q.contiguous(), (h, w), norm), rtol=1e-5, atol=2e-6)
