#!/usr/bin/env python3
"""Per-op microbenchmark of the hot-path kernels at the BASELINE config shapes (profiling aid).

    python tools/kbench.py [--ops recon_tail,in_act,...] [--arch eb4] [--batch 32] [--iters 20]

Each op is timed with CUDA events on the launching stream after warm-up, with a 256 MB buffer written between
iterations (L2 flush).  Prints algorithmic GB/s against MEASURED_PEAKS.json; `bench.py` is the contract
benchmark, this is the inner-loop tool used while tuning one kernel."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def timeit(fn, iters, flush):
    """-> (mean ms between events around the C-ABI calls only, best ms of the same).  The flush is enqueued
    twice before each iteration so the host runs ahead and the op's kernels start back to back."""
    from unidefense_b200 import _lib as L
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    best = 1e9
    for _ in range(iters):
        flush.zero_()
        flush.zero_()
        L.PROFILE = {}
        fn()
        torch.cuda.synchronize()
        t = sum(v[1] for v in L.profile_summary().values())
        L.PROFILE = None
        tot += t
        best = min(best, t)
    return tot / iters, best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", default="recon_tail,in_act,tanh,attention,dyfi,losses")
    ap.add_argument("--arch", default="eb4")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--res", type=int, default=0)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default="", help="in_act: restrict to one plane shape, e.g. 20x192")
    a = ap.parse_args()
    from unidefense_b200 import ops
    _, _, res, nb = bench.ARCH[a.arch]
    res, nb = a.res or res, a.batch or nb
    dev = torch.device("cuda", 0)
    pk, _ = bench.peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    planes, h = bench.decoder_planes(a.arch, res)
    act = "swish" if a.arch == "eb4" else "relu"
    want = set(a.ops.split(","))
    rows = []

    def report(name, ms, best, nbytes):
        gbs = nbytes / (ms * 1e-3) / 1e9
        rows.append((name, ms, best, nbytes / 1e6, gbs, gbs / pk["hbm_gbs"]))
        print(f"{name:34s} {ms * 1e3:9.1f} us (best {best * 1e3:8.1f})  {nbytes / 1e6:9.2f} MB  {gbs:8.1f} GB/s  "
              f"{100 * gbs / pk['hbm_gbs']:5.1f}% of {pk['hbm_gbs']:.0f}", flush=True)

    if "recon_tail" in want:
        dec = torch.tanh(torch.randn(nb, 3, h, h, device=dev)).requires_grad_()
        x = torch.rand(nb, 3, res, res, device=dev) * 2 - 1
        ab = bench.alg_bytes(a.arch, nb, res, nb // 2)
        ms, best = timeit(lambda: ops.recon_tail(dec, x), a.iters, flush)
        report("recon_tail_fwd", ms, best, ab["recon_tail_fwd"])
        rec, sp, fr = ops.recon_tail(dec, x)
        g = torch.zeros(nb, device=dev)
        g[: nb // 2] = 1.0 / (nb // 2)

        def bwd():
            dec.grad = None
            torch.autograd.backward([sp, fr], [0.1 * g, g], retain_graph=True)
        ms, best = timeit(bwd, a.iters, flush)
        report("recon_tail_bwd (real half)", ms, best, ab["recon_tail_bwd"])
    for io, tag, nby in ((torch.float32, "in_act", 4), (torch.bfloat16, "in_act_bf16", 2)):
        if tag not in want:
            continue
        for c, s in sorted(set(planes)):
            if a.only and a.only != f"{c}x{s}":
                continue
            x = torch.randn(nb, c, s, s, device=dev).to(io).requires_grad_()
            gamma = torch.rand(c, device=dev, requires_grad=True)
            beta = torch.randn(c, device=dev, requires_grad=True)
            E = nb * c * s * s
            ms, best = timeit(lambda: ops.in_act(x, gamma, beta, act), a.iters, flush)
            report(f"{tag}_fwd {c}x{s}x{s}", ms, best, 2 * E * nby)
            y = ops.in_act(x, gamma, beta, act)
            gy = torch.randn_like(y)

            def bwd():
                x.grad = None
                y.backward(gy, retain_graph=True)
            ms, best = timeit(bwd, a.iters, flush)
            report(f"{tag}_bwd {c}x{s}x{s}", ms, best, 3 * E * nby)
    if "tanh" in want:
        x = torch.randn(nb, 3, h, h, device=dev, requires_grad=True)
        ms, best = timeit(lambda: ops.tanh(x), a.iters, flush)
        report("tanh_fwd", ms, best, 2 * x.numel() * 4)
    if "attention" in want:
        C, s = {"eb4": (272, -(-res // 32)), "r18": (512, -(-res // 16)), "r50": (2048, -(-res // 32))}[a.arch]
        emb = torch.randn(nb, C, s, s, device=dev, requires_grad=True)
        pred = torch.randn(nb, 3, h, h, device=dev)
        x = torch.rand(nb, 3, res, res, device=dev)
        ms, best = timeit(lambda: ops.attn_prep(pred, x, (s, s)), a.iters, flush)
        report("attn_prep", ms, best, (27 * s * s + 6 * s * (s // 2 + 1)) * 4 * nb)
        ms, best = timeit(lambda: ops.rfft2_cat(emb), a.iters, flush)
        report("rfft2_cat", ms, best, (emb.numel() + 2 * C * s * (s // 2 + 1) * nb) * 4)
        xf = ops.rfft2_cat(emb).detach()
        ms, best = timeit(lambda: ops.irfft2_cat(xf, (s, s)), a.iters, flush)
        report("irfft2_cat", ms, best, (emb.numel() + xf.numel()) * 4)
        smask = torch.rand(nb, 1, s, s, device=dev)
        coef = torch.tensor(0.3, device=dev)
        ms, best = timeit(lambda: ops.attn_fuse(emb, smask, emb, None, coef), a.iters, flush)
        report("attn_fuse_fwd", ms, best, 3 * emb.numel() * 4)
    if "losses" in want:
        labels = torch.tensor([0] * (nb // 2) + [1] * (nb - nb // 2), device=dev)
        for c in ((160, 80, 40) if a.arch == "eb4" else (1024, 256)):
            f = torch.randn(nb, c, device=dev, requires_grad=True)
            ms, best = timeit(lambda: ops.triplet_loss(f, labels), a.iters, flush)
            report(f"triplet c={c}", ms, best, 2 * f.numel() * 4)
        F_ = {"eb4": 1792, "r18": 512, "r50": 2048}[a.arch]
        ea = torch.randn(nb, F_, device=dev, requires_grad=True)
        eb = torch.randn(nb, F_, device=dev)
        ms, best = timeit(lambda: ops.factorization_loss(ea, eb), a.iters, flush)
        report(f"factorization F={F_}", ms, best, 3 * ea.numel() * 4)
    print(json.dumps({"kbench": [{"op": r[0], "ms": r[1], "best_ms": r[2], "alg_mb": r[3], "gbs": r[4], "frac": r[5]}
                                 for r in rows]}))


if __name__ == "__main__":
    main()
