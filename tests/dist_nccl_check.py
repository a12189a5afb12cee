#!/usr/bin/env python3
"""Multi-GPU parity of the hot path under the engines' wrapping (SyncBatchNorm + DDP), real kernels, NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/dist_nccl_check.py

Each rank runs the SyncBN-converted dynamic filters + DDP'd UDR18 hot path on its shard; rank 0 also runs the
same model single-process on the concatenated batch (plain BatchNorm) and compares masks, outputs and the
rank-summed parameter gradients (SURVEY.md §4 'distributed' row)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import procedural as P  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from unidefense_b200.model.modules import FrequencyDynamicFilter, SpatialDynamicFilter, MemoryEfficientSwish

    C, h, w, n_per = 24, 6, 6, 3
    ok = True
    for kind, cls, cin, D in (("freq", FrequencyDynamicFilter, 2 * C, 6), ("spat", SpatialDynamicFilter, C, 3)):
        torch.manual_seed(0)
        ref = cls(C, MemoryEfficientSwish, nn.BatchNorm2d, True, False)
        P.fill_state_dict_(ref, salt=9)
        ref = ref.to(dev).train()
        import copy
        par = nn.SyncBatchNorm.convert_sync_batchnorm(copy.deepcopy(ref)).to(dev).train()
        g = torch.Generator().manual_seed(5)
        X = torch.randn(world * n_per, cin, h, w, generator=g).to(dev)
        Dm = torch.rand(world * n_per, D, h, w, generator=g).to(dev)
        Gm = torch.randn(world * n_per, 1, h, w, generator=g).to(dev)
        Go = torch.randn(world * n_per, cin, h, w, generator=g).to(dev)
        sl = slice(rank * n_per, (rank + 1) * n_per)
        x = X[sl].clone().requires_grad_()
        o = par(x, Dm[sl])
        ((o["mask"] * Gm[sl]).sum() + (o["out"] * Go[sl]).sum()).backward()
        grads = torch.cat([p.grad.flatten() for p in par.parameters()])
        dist.all_reduce(grads)                               # sum over ranks == full-batch gradient
        masks = [torch.empty_like(o["mask"]) for _ in range(world)]
        dist.all_gather(masks, o["mask"].detach().contiguous())
        gxs = [torch.empty_like(x.grad) for _ in range(world)]
        dist.all_gather(gxs, x.grad.contiguous())
        if rank == 0:
            xf = X.clone().requires_grad_()
            of = ref(xf, Dm)
            ((of["mask"] * Gm).sum() + (of["out"] * Go).sum()).backward()
            gref = torch.cat([p.grad.flatten() for p in ref.parameters()])
            try:
                torch.testing.assert_close(torch.cat(masks), of["mask"].detach(), rtol=1e-4, atol=1e-5)
                torch.testing.assert_close(torch.cat(gxs), xf.grad, rtol=2e-4, atol=2e-5 * float(xf.grad.abs().max()))
                torch.testing.assert_close(grads, gref, rtol=2e-4, atol=2e-5 * float(gref.abs().max()))
                torch.testing.assert_close(par.layer1[1].running_var, ref.layer1[1].running_var, rtol=1e-4, atol=1e-6)
                print(f"[dist_nccl_check] {kind} filter: SyncBN x{world} == single-process full batch: OK", flush=True)
            except AssertionError as e:
                ok = False
                print(f"[dist_nccl_check] {kind} filter FAILED: {e}", flush=True)

    # host-sync-free SyncBatchNorm (unidefense_b200/parallel.py) == torch.nn.SyncBatchNorm on the same shards
    from unidefense_b200.parallel import convert_sync_batchnorm
    for cl in (False, True):
        torch.manual_seed(3)
        base = nn.Sequential(nn.Conv2d(6, 16, 3, padding=1, bias=False), nn.BatchNorm2d(16), nn.ReLU(),
                             nn.Conv2d(16, 8, 1, bias=False), nn.BatchNorm2d(8, momentum=0.01, eps=1e-3)).to(dev)
        import copy
        a = nn.SyncBatchNorm.convert_sync_batchnorm(copy.deepcopy(base)).train()
        b = convert_sync_batchnorm(copy.deepcopy(base)).train()
        g = torch.Generator().manual_seed(40 + rank)
        xin = torch.randn(3 + rank, 6, 9, 7, generator=g).to(dev)          # uneven per-rank batch
        gout = torch.randn(3 + rank, 8, 9, 7, generator=g).to(dev)
        if cl:
            a, b = a.to(memory_format=torch.channels_last), b.to(memory_format=torch.channels_last)
            xin = xin.contiguous(memory_format=torch.channels_last)
        outs = []
        for net in (a, b):
            xi = xin.clone().requires_grad_()
            y = net(xi)
            (y * gout).sum().backward()
            outs.append((y.detach(), xi.grad, [p.grad for p in net.parameters()], [bf.clone() for bf in net.buffers()]))
        try:
            torch.testing.assert_close(outs[1][0], outs[0][0], rtol=1e-5, atol=1e-6)
            torch.testing.assert_close(outs[1][1], outs[0][1], rtol=1e-5, atol=1e-6)
            for u, v in zip(outs[1][2], outs[0][2]):
                torch.testing.assert_close(u, v, rtol=1e-5, atol=1e-6)
            for u, v in zip(outs[1][3], outs[0][3]):
                torch.testing.assert_close(u, v, rtol=1e-5, atol=1e-6)
            if rank == 0:
                print(f"[dist_nccl_check] parallel.SyncBatchNorm == torch SyncBatchNorm (channels_last={cl}): OK", flush=True)
        except AssertionError as e:
            ok = False
            print(f"[dist_nccl_check] parallel.SyncBatchNorm FAILED on rank {rank} (channels_last={cl}): {e}", flush=True)

    # whole model under DDP: one step runs, every parameter gets a gradient, ranks agree after the all-reduce
    from unidefense_b200.model import load_model
    torch.manual_seed(0)
    model = load_model("UDR18")(num_classes=2, drop_rate=0.0)
    P.fill_state_dict_(model, salt=5)
    model = convert_sync_batchnorm(model).to(dev).train()
    ddp = nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=False)
    g = torch.Generator().manual_seed(100 + rank)
    x = (torch.rand(4, 3, 96, 96, generator=g) * 2 - 1).to(dev)
    labels = torch.tensor([0, 0, 1, 1], device=dev)
    out = ddp(x)
    from unidefense_b200 import ops
    ld = out["loss_dict"]
    loss = (torch.nn.functional.cross_entropy(out["cls_out"], labels) + 0.1 * ld["freq_mask"].mean()
            + 0.1 * ld["spat_mask"].mean() + 0.1 * sum(ops.triplet_loss(f, labels) for f in ld["triplet"])
            + 0.1 * ld["spatial"][:2].mean() + ld["freq"][:2].mean())
    loss.backward()
    missing = [n for n, p in model.named_parameters() if p.requires_grad and p.grad is None]
    gsum = torch.stack([p.grad.double().abs().sum() for p in model.parameters() if p.grad is not None]).sum()
    both = [torch.zeros_like(gsum) for _ in range(world)]
    dist.all_gather(both, gsum)
    if rank == 0:
        same = all(bool(torch.equal(both[0], b)) for b in both)
        print(f"[dist_nccl_check] DDP UDR18 step: loss {float(loss):.4f}, params without grad: {len(missing)}, "
              f"ranks hold identical reduced grads: {same}", flush=True)
        ok = ok and not missing and same and bool(torch.isfinite(loss))
    # ---- NVLink peer-memory exchange (csrc/ud_comm.cu): gather / reduce vs NCCL, inside a CUDA graph too ----
    from unidefense_b200 import parallel as PAR
    comm = PAR.PeerComm(max_count=9000)
    try:
        for n in (1, 7, 545, 8193):
            v = torch.randn(n, generator=torch.Generator().manual_seed(1000 * n + rank)).to(dev)
            got = comm.gather(v)
            want = [torch.empty_like(v) for _ in range(world)]
            dist.all_gather(want, v)
            assert torch.equal(got, torch.stack(want)), f"gather n={n}"
            red = comm.reduce(v)
            acc = torch.zeros_like(v)
            for t in want:                      # fixed rank order, like the kernel
                acc = acc + t
            assert torch.equal(red, acc), f"reduce n={n}"
        v = torch.randn(1089, device=dev)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                comm.gather(v)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            outg = comm.gather(v * 2.0)
            outr = comm.reduce(v + 1.0)
        for it in range(5):
            v.copy_(torch.randn(1089, generator=torch.Generator().manual_seed(77 * it + rank)).to(dev))
            graph.replay()
            want = [torch.empty_like(v) for _ in range(world)]
            dist.all_gather(want, v)
            assert torch.equal(outg, torch.stack(want) * 2.0), f"graph gather replay {it}"
            acc = torch.zeros_like(v)
            for t in want:
                acc = acc + (t + 1.0)
            assert torch.equal(outr, acc), f"graph reduce replay {it}"
        assert comm.error() == 0
        if rank == 0:
            print(f"[dist_nccl_check] PeerComm gather/reduce == NCCL (eager + 5 CUDA-graph replays) x{world}: OK", flush=True)
    except AssertionError as e:
        ok = False
        print(f"[dist_nccl_check] PeerComm FAILED on rank {rank}: {e}", flush=True)

    # SyncBatchNorm + dynamic filters through the peer exchange == through NCCL; one rank with an EMPTY shard
    PAR.set_default_comm(comm)
    try:
        torch.manual_seed(3)
        base = nn.Sequential(nn.Conv2d(6, 16, 3, padding=1, bias=False), nn.BatchNorm2d(16), nn.ReLU(),
                             nn.Conv2d(16, 8, 1, bias=False), nn.BatchNorm2d(8, momentum=0.01, eps=1e-3)).to(dev)
        import copy
        a = nn.SyncBatchNorm.convert_sync_batchnorm(copy.deepcopy(base)).train()
        b = convert_sync_batchnorm(copy.deepcopy(base)).train()
        for nb in (3 + rank, 0 if rank == world - 1 else 4):            # uneven shards, then an empty last rank
            g = torch.Generator().manual_seed(40 + rank)
            xin = torch.randn(nb, 6, 9, 7, generator=g).to(dev)
            gout = torch.randn(nb, 8, 9, 7, generator=g).to(dev)
            outs = []
            for net in (a, b):
                xi = xin.clone().requires_grad_()
                y = net(xi)
                (y * gout).sum().backward()
                outs.append((y.detach(), torch.zeros_like(xi) if xi.grad is None else xi.grad,
                             [torch.zeros_like(p) if p.grad is None else p.grad.clone()
                                                   for p in net.parameters()], [bf.clone() for bf in net.buffers()]))
                net.zero_grad()
            torch.testing.assert_close(outs[1][0], outs[0][0], rtol=1e-5, atol=1e-6)
            torch.testing.assert_close(outs[1][1], outs[0][1], rtol=1e-5, atol=1e-6)
            for u, v2 in zip(outs[1][2], outs[0][2]):
                torch.testing.assert_close(u, v2, rtol=1e-5, atol=1e-6)
            for u, v2 in zip(outs[1][3], outs[0][3]):
                assert bool(torch.isfinite(u.float()).all()), "non-finite running statistics"
                torch.testing.assert_close(u, v2, rtol=1e-5, atol=1e-6)
        if rank == 0:
            print("[dist_nccl_check] parallel.SyncBatchNorm over PeerComm == torch SyncBatchNorm (incl. an empty rank): OK",
                  flush=True)
    except AssertionError as e:
        ok = False
        print(f"[dist_nccl_check] SyncBatchNorm over PeerComm FAILED on rank {rank}: {e}", flush=True)

    # FlatGradients: one all-reduce of the flat buffer == DDP's averaged gradients
    try:
        torch.manual_seed(0)
        m2 = load_model("UDR18")(num_classes=2, drop_rate=0.0)
        P.fill_state_dict_(m2, salt=5)
        m2 = convert_sync_batchnorm(m2).to(dev).train()
        fg = PAR.FlatGradients(m2.parameters())
        fg.zero()
        out2 = m2(x)
        ld2 = out2["loss_dict"]
        loss2 = (torch.nn.functional.cross_entropy(out2["cls_out"], labels) + 0.1 * ld2["freq_mask"].mean()
                 + 0.1 * ld2["spat_mask"].mean() + 0.1 * sum(ops.triplet_loss(f, labels) for f in ld2["triplet"])
                 + 0.1 * ld2["spatial"][:2].mean() + ld2["freq"][:2].mean())
        loss2.backward()
        fg.all_reduce()
        worst = 0.0
        for (n1, p1), (n2, p2) in zip(model.named_parameters(), m2.named_parameters()):
            if p1.grad is None:
                continue
            d = float((p1.grad - p2.grad).abs().max()) / (float(p1.grad.abs().max()) + 1e-12)
            worst = max(worst, d)
        assert worst < 2e-3, f"FlatGradients vs DDP: worst relative gradient difference {worst:.2e}"
        if rank == 0:
            print(f"[dist_nccl_check] FlatGradients (one flat all-reduce, peer-memory SyncBN) == DDP gradients "
                  f"(worst rel diff {worst:.1e}): OK", flush=True)
    except AssertionError as e:
        ok = False
        print(f"[dist_nccl_check] FlatGradients FAILED on rank {rank}: {e}", flush=True)
    PAR.set_default_comm(None)

    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
