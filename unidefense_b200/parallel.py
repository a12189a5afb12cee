"""Data-parallel plumbing: a SyncBatchNorm that never synchronises the host.

torch.nn.SyncBatchNorm's forward filters out zero-count ranks with a boolean-mask index
(`count_all[mask]`), which forces a device->host synchronisation in EVERY layer and step unless the
stream is being captured.  UniDefense's engines convert all 99 (EfficientNet-B4) BatchNorms
(engine/forgery_engine.py:142), so the host can never run ahead of the GPU: the multi-GPU step grows by
a fixed ~56 ms independent of the number of ranks.  This module keeps the same math --
the same ATen primitives (batch_norm_stats, batch_norm_gather_stats_with_counts, batch_norm_elemt,
batch_norm_backward_reduce, batch_norm_backward_elemt) and the same two collectives per layer -- but never
reads the counts on the host: zero-count ranks contribute zero weight to the gathered statistics anyway.

    model = unidefense_b200.parallel.convert_sync_batchnorm(model)      # instead of nn.SyncBatchNorm.convert_...

The class subclasses nn.SyncBatchNorm, so isinstance checks (ours in model/modules.py, DDP's, user code) and
state_dict keys are unchanged.
"""
import ctypes

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F


class PeerComm:
    """Small-vector exchange between the ranks of ONE NVLink/NVSwitch box over CUDA-IPC peer memory
    (csrc/ud_comm.cu): `gather(v)` -> [world, n], `reduce(v)` -> sum over ranks, each ONE single-CTA kernel per
    rank (remote stores + release/acquire flags), no NCCL call, no host synchronisation, CUDA-graph capturable.
    Built once per process from an initialised process group (the IPC handles travel through it)."""

    def __init__(self, group=None, max_count=16384):
        from . import _lib as L
        self.L = L
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.max_count = int(max_count)
        lib = L.lib()
        nbytes = lib.ud_comm_buffer_bytes(self.world, self.max_count)
        if nbytes == 0:
            raise RuntimeError(f"PeerComm: world size {self.world} unsupported (<= 16)")
        own = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        L.check(lib.ud_comm_alloc(nbytes, ctypes.byref(own), handle), "comm_alloc")
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw), group=self.group)
        peers = (ctypes.c_void_p * self.world)()
        for q, hq in enumerate(handles):
            if q == self.rank:
                peers[q] = own.value
            else:
                ptr = ctypes.c_void_p()
                L.check(lib.ud_comm_open(ctypes.create_string_buffer(hq, 64), ctypes.byref(ptr)), "comm_open")
                peers[q] = ptr.value
        comm = ctypes.c_void_p()
        L.check(lib.ud_comm_create(peers, self.rank, self.world, self.max_count, ctypes.byref(comm)), "comm_create")
        self.handle = comm
        dist.barrier(group=self.group)          # every buffer is mapped everywhere before the first gather

    def _run(self, v, reduce):
        L = self.L
        v = v.contiguous()
        L.require_cuda_f32(v)
        n = v.numel()
        out = torch.empty(n if reduce else (self.world, n), device=v.device, dtype=torch.float32)
        L.check(L.lib().ud_comm_gather(self.handle, L.ptr(v), L.ptr(out), n, int(reduce), L.stream()), "comm_gather")
        return out

    def gather(self, v):
        """[n] fp32 on every rank -> [world, n]."""
        return self._run(v, False)

    def reduce(self, v):
        """[n] (any shape) fp32 -> elementwise sum over ranks, same shape (fixed rank order: deterministic)."""
        return self._run(v, True).view(v.shape)

    def error(self):
        return self.L.lib().ud_comm_error(self.handle)


_default_comm = None


def set_default_comm(comm):
    """Make `comm` the exchange used by every converted SyncBatchNorm and by the dynamic filters' BatchNorms
    (None: back to torch.distributed collectives)."""
    global _default_comm
    _default_comm = comm


def default_comm():
    return _default_comm


class FlatGradients:
    """Batch-sharded data parallelism without DDP's reducer: every parameter's .grad is a view into ONE flat fp32
    buffer, so the gradient exchange of a step is a single NCCL all-reduce (513 MB for UDEB4 at ~0.7 TB/s bus
    bandwidth is ~1.3 ms of a 77 ms step: no bucketing/overlap machinery needed) and the whole step -- forward,
    backward, all-reduce, optimizer -- captures into one CUDA graph.  Semantics = DDP's: gradients are averaged."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        off = 0
        for p in self.params:
            n = p.numel()
            # same strides as the parameter (channels_last weights included): autograd accumulates in place
            p.grad = torch.as_strided(self.flat, p.shape, p.stride(), off)
            off += n

    def zero(self):
        self.flat.zero_()

    def all_reduce(self):
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            if dist.get_backend(self.group) == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
            else:                                   # gloo (CPU tests) has no AVG
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
                self.flat /= float(dist.get_world_size(self.group))


class _SyncBNFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum, group, world_size):
        if not (x.is_contiguous(memory_format=torch.channels_last) or x.is_contiguous(memory_format=torch.channels_last_3d)):
            x = x.contiguous()
        if weight is not None:
            weight = weight.contiguous()
        C = x.shape[1]
        per_channel = x.numel() // C
        if x.numel() > 0:
            mean, invstd = torch.batch_norm_stats(x, eps)
            packed = torch.cat([mean, invstd, mean.new_full((1,), float(per_channel))])
        else:
            packed = torch.zeros(2 * C + 1, dtype=torch.float32, device=x.device)
        comm = _default_comm
        gathered = None
        if comm is not None and 2 * C + 1 <= comm.max_count:
            gathered = comm.gather(packed.float())            # one peer-memory kernel instead of an NCCL all_gather
        elif group._get_backend_name() != "gloo":
            gathered = torch.empty(world_size, 2 * C + 1, dtype=packed.dtype, device=packed.device)
            dist.all_gather_into_tensor(gathered.view(1, -1), packed, group)
        else:
            parts = [torch.empty_like(packed) for _ in range(world_size)]
            dist.all_gather(parts, packed, group)
            gathered = torch.stack(parts)
        mean_all, invstd_all, count_all = gathered[:, :C], gathered[:, C:2 * C], gathered[:, 2 * C]
        # a rank with an empty input carries (0, 0, count 0): neutralise its invstd on the device (1/invstd^2 would
        # be inf and inf*0 = NaN inside batch_norm_gather_stats_with_counts) -- still no host synchronisation
        invstd_all = torch.where(count_all[:, None] > 0, invstd_all, torch.ones_like(invstd_all))
        counts = count_all.reshape(-1)
        if running_mean is not None and counts.dtype != running_mean.dtype:
            counts = counts.to(running_mean.dtype)
        # ranks with count 0 carry zero weight in the merge: no need to drop them (and no host sync)
        mean, invstd = torch.batch_norm_gather_stats_with_counts(x, mean_all.contiguous(), invstd_all.contiguous(),
                                                                 running_mean, running_var, momentum, eps, counts)
        ctx.save_for_backward(x, weight, mean, invstd, count_all.reshape(-1, 1).to(torch.int32))
        ctx.group = group
        ctx.comm = comm if (comm is not None and 2 * C <= comm.max_count) else None
        if x.numel() == 0:
            return torch.empty_like(x)
        return torch.batch_norm_elemt(x, weight, bias, mean, invstd, eps)

    @staticmethod
    def backward(ctx, gy):
        if not (gy.is_contiguous(memory_format=torch.channels_last) or gy.is_contiguous(memory_format=torch.channels_last_3d)):
            gy = gy.contiguous()
        x, weight, mean, invstd, counts = ctx.saved_tensors
        gx = gw = gb = None
        if x.numel() > 0:
            sum_dy, sum_dy_xmu, gw, gb = torch.batch_norm_backward_reduce(
                gy, x, mean, invstd, weight, ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2])
            if ctx.needs_input_grad[0]:
                C = sum_dy.shape[0]
                both = torch.cat([sum_dy, sum_dy_xmu])
                if ctx.comm is not None:
                    both = ctx.comm.reduce(both.float())
                else:
                    dist.all_reduce(both, op=dist.ReduceOp.SUM, group=ctx.group)
                sum_dy, sum_dy_xmu = both[:C], both[C:]
                w = weight if weight is None or weight.dtype == mean.dtype else weight.to(mean.dtype)
                gx = torch.batch_norm_backward_elemt(gy, x, mean, invstd, w, sum_dy, sum_dy_xmu, counts)
        else:
            C = x.shape[1]
            if ctx.needs_input_grad[0]:
                both = torch.zeros(2 * C, dtype=mean.dtype, device=x.device)
                if ctx.comm is not None:
                    ctx.comm.reduce(both)                                         # keep the collective order
                else:
                    dist.all_reduce(both, op=dist.ReduceOp.SUM, group=ctx.group)     # keep the collective order
                gx = torch.zeros_like(x)
        if weight is None or not ctx.needs_input_grad[1]:
            gw = None
        if weight is None or not ctx.needs_input_grad[2]:
            gb = None
        return gx, gw, gb, None, None, None, None, None, None


class SyncBatchNorm(nn.SyncBatchNorm):
    """nn.SyncBatchNorm without the per-layer host synchronisation (see module docstring)."""

    def forward(self, input):
        self._check_input_dim(input)
        self._check_non_zero_input_channels(input)
        if self.momentum is None:
            factor = 0.0
        else:
            factor = self.momentum
        if self.training and self.track_running_stats:
            self.num_batches_tracked.add_(1)
            if self.momentum is None:
                factor = 1.0 / float(self.num_batches_tracked)          # cumulative average (host read, as in torch)
        bn_training = self.training or (self.running_mean is None and self.running_var is None)
        running_mean = self.running_mean if not self.training or self.track_running_stats else None
        running_var = self.running_var if not self.training or self.track_running_stats else None
        need_sync = bn_training and self.training and dist.is_available() and dist.is_initialized()
        group, world = None, 1
        if need_sync:
            if input.device.type not in ("cuda", "xpu", "hpu"):
                raise ValueError("SyncBatchNorm expected input tensor to be on GPU")
            group = self.process_group if self.process_group else dist.group.WORLD
            world = dist.get_world_size(group)
            need_sync = world > 1
        if not need_sync:
            return F.batch_norm(input, running_mean, running_var, self.weight, self.bias, bn_training, factor, self.eps)
        return _SyncBNFunction.apply(input, self.weight, self.bias, running_mean, running_var, self.eps, factor, group, world)


def convert_sync_batchnorm(module, process_group=None):
    """Drop-in for torch.nn.SyncBatchNorm.convert_sync_batchnorm (engine/forgery_engine.py:142)."""
    out = module
    if isinstance(module, nn.modules.batchnorm._BatchNorm) and not isinstance(module, SyncBatchNorm):
        out = SyncBatchNorm(module.num_features, module.eps, module.momentum, module.affine, module.track_running_stats,
                            process_group)
        if module.affine:
            with torch.no_grad():
                out.weight = module.weight
                out.bias = module.bias
        out.running_mean = module.running_mean
        out.running_var = module.running_var
        out.num_batches_tracked = module.num_batches_tracked
        out.training = module.training
        if hasattr(module, "qconfig"):
            out.qconfig = module.qconfig
    for name, child in module.named_children():
        out.add_module(name, convert_sync_batchnorm(child, process_group))
    return out


# ---- harness-side synchronisation -----------------------------------------------------------
def reduce_tensor(t, group=None):
    """Drop-in for the reference's utils/misc.py:18-22 (`all_reduce` then divide by the world size)."""
    rt = t.detach().clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(rt, group=group)
        rt /= float(dist.get_world_size(group))
    return rt


def reduce_scalars(values, group=None):
    """The logging tail of the reference's training loops (engine/forgery_engine.py:279-287, ocim_engine.py:271-279,
    uniattack_engine.py:333-341) calls `reduce_tensor(v).item()` once per logged loss and once for the accuracy: ~10
    latency-bound all-reduces, each followed by a host synchronisation, every step.  This packs every scalar of
    `values` (a dict name -> 0-d / 1-element tensor or float) into ONE tensor, averages it over the ranks with ONE
    all-reduce and reads it back with ONE device-to-host copy.  Returns {name: float} with exactly the numbers the
    per-key calls produce (same fp32 sum over ranks, same division)."""
    names = list(values)
    if not names:
        return {}
    dev = next((v.device for v in values.values() if torch.is_tensor(v)), torch.device("cpu"))
    packed = torch.stack([(v.detach().reshape(-1)[0] if torch.is_tensor(v) else torch.tensor(float(v))).to(dev, torch.float32)
                          for v in values.values()])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(packed, group=group)
        packed /= float(dist.get_world_size(group))
    host = packed.tolist()                     # the step's only host synchronisation for logging
    return dict(zip(names, host))


def logged_losses(out_dict):
    """The keys the reference loops log: every entry of the engine's return dict whose name contains "loss"."""
    return {k: v for k, v in out_dict.items() if "loss" in k}
