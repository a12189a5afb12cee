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
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F


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
        gathered = torch.empty(world_size, 2 * C + 1, dtype=packed.dtype, device=packed.device)
        if group._get_backend_name() != "gloo":
            dist.all_gather_into_tensor(gathered.view(1, -1), packed, group)
        else:
            parts = [torch.empty_like(packed) for _ in range(world_size)]
            dist.all_gather(parts, packed, group)
            gathered = torch.stack(parts)
        mean_all, invstd_all, count_all = gathered[:, :C], gathered[:, C:2 * C], gathered[:, 2 * C]
        counts = count_all.reshape(-1)
        if running_mean is not None and counts.dtype != running_mean.dtype:
            counts = counts.to(running_mean.dtype)
        # ranks with count 0 carry zero weight in the merge: no need to drop them (and no host sync)
        mean, invstd = torch.batch_norm_gather_stats_with_counts(x, mean_all.contiguous(), invstd_all.contiguous(),
                                                                 running_mean, running_var, momentum, eps, counts)
        ctx.save_for_backward(x, weight, mean, invstd, count_all.reshape(-1, 1).to(torch.int32))
        ctx.group = group
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
                dist.all_reduce(both, op=dist.ReduceOp.SUM, group=ctx.group)
                sum_dy, sum_dy_xmu = both[:C], both[C:]
                w = weight if weight is None or weight.dtype == mean.dtype else weight.to(mean.dtype)
                gx = torch.batch_norm_backward_elemt(gy, x, mean, invstd, w, sum_dy, sum_dy_xmu, counts)
        else:
            C = x.shape[1]
            if ctx.needs_input_grad[0]:
                both = torch.zeros(2 * C, dtype=mean.dtype, device=x.device)
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
