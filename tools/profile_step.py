#!/usr/bin/env python3
"""Kernel-level breakdown of one bench.py training step with torch.profiler (profiling aid, not a benchmark).
    python tools/profile_step.py [--arch eb4] [--batch 32] > profiles/step_breakdown.txt"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="eb4")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--rows", type=int, default=45)
    a = ap.parse_args()
    from unidefense_b200 import ops
    from unidefense_b200.model import load_model
    name, kw, res, nb = bench.ARCH[a.arch]
    nb = a.batch or nb
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    torch.backends.cudnn.benchmark = True
    model = bench.init_live(load_model(name)(**kw)).to(dev).train()
    for part in ("backbone", "extractor", "emb_block1", "emb_block2"):
        if hasattr(model, part):
            getattr(model, part).to(memory_format=torch.channels_last)
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=5e-6, amsgrad=True,
                            fused=True)
    x, labels = bench.synth(nb, res, 0, dev)
    nr = nb // 2
    lam = bench.LAMBDAS

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=a.dtype == "bf16"):
            out = model(x)
        ld = out["loss_dict"]
        tri = sum(ops.triplet_loss(f, labels) for f in ld["triplet"])
        loss = (F.cross_entropy(out["cls_out"].float(), labels) + lam["mask"] * ld["freq_mask"].mean()
                + lam["mask"] * ld["spat_mask"].mean() + lam["triplet"] * tri + lam["recons"] * ld["spatial"][:nr].mean()
                + lam["freq"] * ld["freq"][:nr].mean())
        loss.backward()
        opt.step()

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=a.rows, max_name_column_width=90))


if __name__ == "__main__":
    main()
