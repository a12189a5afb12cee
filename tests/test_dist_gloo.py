"""N>1 host logic on CPU with world_size-2 gloo groups: the SyncBatchNorm statistics exchange
of the dynamic filters (all-gather of [mean, M2, count] + Chan merge forward, all-reduce of [sum dz, sum dz*xh]
backward) and bench.py's rank plumbing.  The CUDA statistics kernel is replaced by a torch stand-in here
(a test double, the product path has no CPU fallback); `tests/dist_nccl_check.py` runs the real kernels on 2 GPUs."""
import os
import socket
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cpu_local_stats(proj):
    mean = proj.mean(dim=(0, 2, 3))
    m2 = ((proj - mean.view(1, -1, 1, 1)) ** 2).sum(dim=(0, 2, 3))
    return mean, m2


def _worker(rank, world, port, shards, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from unidefense_b200 import ops
        from unidefense_b200.model import modules as M
        ops.bn_local_stats = _cpu_local_stats           # test double for the CUDA kernel
        bn = nn.SyncBatchNorm(shards[0].shape[1])
        bn.train()
        proj = shards[rank]
        mean, rstd, count, reduce_fn = M.bn_forward_stats(bn, proj)
        sums = torch.stack([proj.sum(dim=(0, 2, 3)), (proj ** 2).sum(dim=(0, 2, 3))])
        red = reduce_fn(sums.clone())
        # uneven shards: rank 1 owns more samples than rank 0
        out[rank] = dict(mean=mean.clone(), rstd=rstd.clone(), count=float(count), red=red.clone(),
                         running_mean=bn.running_mean.clone(), running_var=bn.running_var.clone(),
                         tracked=int(bn.num_batches_tracked))
    finally:
        dist.destroy_process_group()


def test_syncbn_statistics_exchange_two_ranks():
    g = torch.Generator().manual_seed(0)
    full = torch.randn(5, 6, 4, 3, generator=g) * 2.0 + 0.5
    shards = [full[:2].contiguous(), full[2:].contiguous()]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), shards, out), nprocs=2, join=True)
    ref_bn = nn.BatchNorm2d(6)
    ref_bn.train()
    ref_bn(full)
    gmean = full.mean(dim=(0, 2, 3))
    gvar = full.var(dim=(0, 2, 3), unbiased=False)
    for r in (0, 1):
        o = out[r]
        torch.testing.assert_close(o["mean"], gmean, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(o["rstd"], torch.rsqrt(gvar + 1e-5), rtol=1e-5, atol=1e-6)
        assert o["count"] == 5 * 4 * 3
        torch.testing.assert_close(o["running_mean"], ref_bn.running_mean, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(o["running_var"], ref_bn.running_var, rtol=1e-5, atol=1e-6)
        assert o["tracked"] == 1
        want = torch.stack([full.sum(dim=(0, 2, 3)), (full ** 2).sum(dim=(0, 2, 3))])
        torch.testing.assert_close(o["red"], want, rtol=1e-5, atol=1e-5)


def test_local_bn_needs_no_group():
    from unidefense_b200.model import modules as M
    from unidefense_b200 import ops
    old = ops.bn_local_stats
    ops.bn_local_stats = _cpu_local_stats
    try:
        bn = nn.BatchNorm2d(3)
        x = torch.randn(4, 3, 2, 2)
        mean, rstd, count, reduce_fn = M.bn_forward_stats(bn, x)
        assert reduce_fn is None and count == 16
        torch.testing.assert_close(mean, x.mean(dim=(0, 2, 3)))
        bn.eval()
        mean2, rstd2, count2, _ = M.bn_forward_stats(bn, x)
        assert count2 == 0 and mean2 is bn.running_mean
    finally:
        ops.bn_local_stats = old


def test_bench_reference_arm_only_rank0_prints():
    """Under torchrun the reference arm runs on rank 0 alone; the other ranks exit 0 without work."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def _scalars_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from unidefense_b200 import parallel as PAR
        g = torch.Generator().manual_seed(100 + rank)
        ret = {"total_loss": torch.rand((), generator=g), "cls_loss": torch.rand(1, generator=g), "cls_out": torch.rand(4, 2),
               "real_rec_loss": torch.rand((), generator=g) * 3, "fac_loss": 0.25 * (rank + 1)}
        acc = torch.rand((), generator=g)
        # the reference's way: one all-reduce + one .item() per logged key (engine/forgery_engine.py:279-287)
        ref = {}
        for k, v in ret.items():
            if "loss" in k:
                t = v if torch.is_tensor(v) else torch.tensor(v)
                rt = t.clone()
                dist.all_reduce(rt)
                rt /= float(dist.get_world_size())
                ref[k] = rt.reshape(-1)[0].item()
        rt = acc.clone()
        dist.all_reduce(rt)
        ref["acc"] = (rt / float(world)).item()
        packed = PAR.reduce_scalars({**PAR.logged_losses(ret), "acc": acc})
        single = PAR.reduce_tensor(ret["total_loss"]).item()
        out[rank] = dict(ref=ref, packed=packed, single=single)
    finally:
        dist.destroy_process_group()


def test_packed_logging_reduction_equals_per_key_reduce_tensor():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_scalars_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    for r in (0, 1):
        o = out[r]
        assert list(o["packed"]) == ["total_loss", "cls_loss", "real_rec_loss", "fac_loss", "acc"]      # "cls_out" is not a loss
        for k, v in o["ref"].items():
            assert o["packed"][k] == pytest.approx(v, rel=1e-6, abs=1e-7), k
        assert o["single"] == pytest.approx(o["ref"]["total_loss"], rel=1e-6)
    assert out[0]["packed"] == out[1]["packed"]


def test_packed_reduction_without_a_process_group():
    from unidefense_b200 import parallel as PAR
    got = PAR.reduce_scalars({"a_loss": torch.tensor(1.5), "b": 2.0})
    assert got == {"a_loss": 1.5, "b": 2.0} and PAR.reduce_scalars({}) == {}
    assert float(PAR.reduce_tensor(torch.tensor(3.0))) == 3.0


def _flat_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from unidefense_b200 import parallel as PAR
        torch.manual_seed(0)                                  # same weights on every rank
        net = nn.Sequential(nn.Conv2d(3, 4, 3, padding=1), nn.ReLU(), nn.Conv2d(4, 2, 1, bias=False))
        net[0].weight.data = net[0].weight.data.contiguous(memory_format=torch.channels_last)     # strided .grad views
        net[2].weight.requires_grad_(True)
        frozen = nn.Parameter(torch.ones(3), requires_grad=False)
        fg = PAR.FlatGradients(list(net.parameters()) + [frozen])
        assert fg.flat.numel() == sum(p.numel() for p in net.parameters()) and frozen.grad is None
        g = torch.Generator().manual_seed(10 + rank)          # a different shard per rank
        steps = []
        for it in range(2):
            fg.zero()
            x = torch.randn(2 + rank, 3, 5, 5, generator=g)
            net(x).square().mean().backward()                 # autograd accumulates straight into the flat buffer
            local = [p.grad.clone() for p in net.parameters()]
            fg.all_reduce()
            steps.append(dict(local=local, reduced=[p.grad.clone() for p in net.parameters()], flat=fg.flat.clone()))
        out[rank] = steps
    finally:
        dist.destroy_process_group()


def test_flat_gradients_average_over_two_ranks():
    """FlatGradients == DDP's semantics (mean of the per-rank gradients), one all-reduce of one buffer; .grad views keep
    the parameters' strides (channels_last weights) and survive zero() / repeated steps."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_flat_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    for it in range(2):
        a, b = out[0][it], out[1][it]
        for la, lb, ra, rb in zip(a["local"], b["local"], a["reduced"], b["reduced"]):
            torch.testing.assert_close(ra, (la + lb) / 2, rtol=1e-6, atol=1e-7)
            assert torch.equal(ra, rb)
        assert torch.equal(a["flat"], b["flat"])
    assert out[0][0]["reduced"][0].is_contiguous(memory_format=torch.channels_last)
