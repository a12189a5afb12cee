#!/usr/bin/env python3
"""Times the tcgen05 filter projections (csrc/ud_proj.cu) against the library convolution torch would run for
model/modules.py:82-85 / :111-114 on the same B200: cuDNN fp32 (TF32 on = torch default, TF32 off) and bf16
channels-last.  CUDA events, L2 flushed between iterations.  Prints one JSON line per shape.

    python tools/proj_bench.py [--iters 20]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

# (name, N, Cin, Cout, H, W, k): attention projections of the three models at their config batch
SHAPES = [("eb4_freq_1x1", 32, 544, 544, 12, 7, 1), ("eb4_spat_3x3", 32, 272, 272, 12, 12, 3),
          ("r18_freq_1x1", 32, 1024, 1024, 16, 9, 1), ("r18_spat_3x3", 32, 512, 512, 16, 16, 3),
          ("r50_freq_1x1", 64, 4096, 4096, 8, 5, 1), ("r50_spat_3x3", 64, 2048, 2048, 8, 8, 3)]


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    best = 1e9
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    from unidefense_b200 import _lib as L
    from unidefense_b200 import ops
    dev = torch.device("cuda")
    torch.backends.cudnn.benchmark = True
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    bf16_peak = pk.get("bf16_tflops", 1590.0)
    rows = []
    for name, N, Cin, Cout, H, W, k in SHAPES:
        if args.only and args.only not in name:
            continue
        g = torch.Generator().manual_seed(1)
        x = torch.randn(N, Cin, H, W, generator=g).to(dev)
        w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(dev)
        x_cl = x.contiguous(memory_format=torch.channels_last)
        flops = 2.0 * N * H * W * Cin * Cout * k * k
        lib = L.lib()
        # operands prepared once (weights change once per step, activations come from our producer kernels)
        x_hi = torch.empty(N, H, W, Cin, device=dev)
        x_lo = torch.empty_like(x_hi)
        w_hi = torch.empty(Cout, k * k, Cin, device=dev)
        w_lo = torch.empty_like(w_hi)
        L.check(lib.ud_proj_prep_x(L.ptr(x), L.ptr(x_hi), L.ptr(x_lo), N, Cin, H * W, L.stream()), "prep_x")
        L.check(lib.ud_proj_prep_w(L.ptr(w), L.ptr(w_hi), L.ptr(w_lo), Cout, Cin, k * k, L.stream()), "prep_w")
        x_raw = x_cl.permute(0, 2, 3, 1).contiguous()
        w_raw = w.permute(0, 2, 3, 1).contiguous()
        y = torch.empty(N, Cout, H, W, device=dev)
        tiles = lib.ud_proj_m_tiles(N, H, W, k)
        pm = torch.empty(tiles, Cout, device=dev)
        p2 = torch.empty_like(pm)
        pc = torch.empty(tiles, device=dev)

        def gemm_tf32():
            L.check(lib.ud_proj_fwd(L.ptr(x_raw), None, L.ptr(w_raw), None, L.ptr(y), L.ptr(pm), L.ptr(p2), L.ptr(pc),
                                    N, H, W, Cin, Cout, k, L.stream()), "proj")

        def gemm_3x():
            L.check(lib.ud_proj_fwd(L.ptr(x_hi), L.ptr(x_lo), L.ptr(w_hi), L.ptr(w_lo), L.ptr(y), L.ptr(pm), L.ptr(p2),
                                    L.ptr(pc), N, H, W, Cin, Cout, k, L.stream()), "proj")

        def full_tf32():
            ops.proj_conv(x, w, "tf32")

        def full_3x():
            ops.proj_conv(x, w, "3xtf32")

        def cudnn(tf32, inp, wt):
            def f():
                torch.backends.cudnn.allow_tf32 = tf32
                out = F.conv2d(inp, wt, None, 1, k // 2)
                ops.bn_local_stats(out)              # the statistics pass our epilogue replaces
            return f

        xb, wb = x_cl.bfloat16(), w.bfloat16().contiguous(memory_format=torch.channels_last)

        def cudnn_bf16():
            out = F.conv2d(xb, wb, None, 1, k // 2)
            ops.bn_local_stats(out.float())

        r = {"shape": name, "N": N, "Cin": Cin, "Cout": Cout, "HW": [H, W], "k": k, "gflop": round(flops / 1e9, 2)}
        for key, fn in [("ours_tf32_gemm", gemm_tf32), ("ours_3xtf32_gemm", gemm_3x), ("ours_tf32_op", full_tf32),
                        ("ours_3xtf32_op", full_3x), ("cudnn_tf32+stats", cudnn(True, x, w)),
                        ("cudnn_fp32+stats", cudnn(False, x, w)), ("cudnn_tf32_cl+stats", cudnn(True, x_cl, w)),
                        ("cudnn_bf16_cl+stats", cudnn_bf16)]:
            ms, best = timeit(fn, args.iters, flush)
            r[key] = {"us": round(ms * 1e3, 1), "best_us": round(best * 1e3, 1), "tflops": round(flops / (best * 1e-3) / 1e12, 1)}
        # numerical check against fp64 while we are here
        ref = F.conv2d(x.double(), w.double(), None, 1, k // 2)
        gemm_tf32()
        r["err_tf32"] = float((y.double() - ref).abs().max() / ref.abs().max())
        gemm_3x()
        r["err_3xtf32"] = float((y.double() - ref).abs().max() / ref.abs().max())
        torch.backends.cudnn.allow_tf32 = True
        r["err_cudnn_tf32"] = float((F.conv2d(x, w, None, 1, k // 2).double() - ref).abs().max() / ref.abs().max())
        r["tensor_pipe_frac_tf32_of_bf16_peak"] = round(r["ours_tf32_gemm"]["tflops"] / bf16_peak, 4)
        rows.append(r)
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
