"""Synchronised BatchNorm whose statistics exchange runs on libdmlb's peer communicator.

The reference offers `register_model(..., sync_bn=True)` (pipeline.py:60,70-71), which converts every BatchNorm layer to
`torch.nn.SyncBatchNorm`: per layer and step, an all_gather of `[mean, invstd, count]` (2C + 1 floats per rank) in the
forward pass and an all_reduce of `[sum_dy, sum_dy_xmu]` (2C floats) in the backward pass — tiny, latency-bound NCCL
collectives (tens of microseconds each on NVSwitch), two per BatchNorm layer per step.

`PeerSyncBatchNorm` keeps torch's arithmetic (the same ATen building blocks: batch_norm_stats ->
batch_norm_gather_stats_with_counts -> batch_norm_elemt, and batch_norm_backward_reduce -> batch_norm_backward_elemt) and
replaces the two collectives with ONE libdmlb kernel each: `dmlb_comm_allreduce` on a small fp32 buffer, which for messages
of this size is the LL protocol (data and flag pushed together into every peer's arena: no barrier, ~6-9 us at 8 GPUs).
  forward   every rank writes its `[mean, invstd, count]` row into a zeroed `[W, 2C + 1]` matrix and the matrix is
            sum-all-reduced: x + 0 is exact, so every rank ends up with exactly the gathered rows torch's all_gather gives
  backward  `[sum_dy, sum_dy_xmu]` is sum-all-reduced in rank order (torch: all_reduce SUM)
Both kernels are CUDA-graph capturable (sequence number in device memory), so a model with SyncBN still runs as a captured
step.  One communicator (its own arena) serves all layers of a pipeline; it is driven on the compute stream.

`convert(module, comm)` mirrors `torch.nn.SyncBatchNorm.convert_sync_batchnorm`.
"""
import torch

from . import _native as N

WIRE_F32 = N.WIRE_F32


def _allreduce_sum_(comm, buf):
    """In-place rank-ordered fp32 sum of the flat fp32 CUDA tensor `buf` over the communicator's ranks."""
    lib = N.cuda_lib(buf.device.index)
    N.check(lib.dmlb_comm_allreduce(comm.handle, buf.data_ptr(), buf.numel(), WIRE_F32, 1.0, None, 0, None, N.stream_ptr()),
            'comm_allreduce(syncbn)')
    return buf


class _PeerSyncBN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum, comm):
        if not (x.is_contiguous(memory_format=torch.channels_last) or x.is_contiguous(memory_format=torch.channels_last_3d)):
            x = x.contiguous()
        if weight is not None:
            weight = weight.contiguous()
        channels = x.shape[1]
        per_channel = x.numel() // channels
        world, rank = comm.world, comm.rank
        mean, invstd = torch.batch_norm_stats(x, eps)
        # [W, 2C + 1] with only this rank's row filled: the sum over ranks IS the all_gather
        rows = torch.zeros(world, 2 * channels + 1, dtype=torch.float32, device=x.device)
        mine = rows[rank]
        mine[:channels].copy_(mean)
        mine[channels:2 * channels].copy_(invstd)
        mine[2 * channels].fill_(float(per_channel))
        _allreduce_sum_(comm, rows.view(-1))
        mean_all, invstd_all, count_all = rows[:, :channels], rows[:, channels:2 * channels], rows[:, 2 * channels]
        counts = count_all.to(running_mean.dtype) if running_mean is not None else count_all
        mean, invstd = torch.batch_norm_gather_stats_with_counts(x, mean_all, invstd_all, running_mean, running_var,
                                                                 momentum, eps, counts.contiguous())
        ctx.save_for_backward(x, weight, mean, invstd, count_all.to(torch.int32))
        ctx.comm = comm
        return torch.batch_norm_elemt(x, weight, bias, mean, invstd, eps)

    @staticmethod
    def backward(ctx, grad_out):
        if not (grad_out.is_contiguous(memory_format=torch.channels_last) or
                grad_out.is_contiguous(memory_format=torch.channels_last_3d)):
            grad_out = grad_out.contiguous()
        x, weight, mean, invstd, count = ctx.saved_tensors
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        sum_dy, sum_dy_xmu, grad_weight, grad_bias = torch.batch_norm_backward_reduce(grad_out, x, mean, invstd, weight,
                                                                                       need_x, need_w, need_b)
        grad_x = None
        if need_x:
            channels = sum_dy.shape[0]
            both = torch.cat([sum_dy, sum_dy_xmu]).to(torch.float32).contiguous()
            _allreduce_sum_(ctx.comm, both)
            sum_dy, sum_dy_xmu = both[:channels], both[channels:]
            if weight is not None and weight.dtype != mean.dtype:
                weight = weight.to(mean.dtype)
            grad_x = torch.batch_norm_backward_elemt(grad_out, x, mean, invstd, weight, sum_dy, sum_dy_xmu, count)
        return grad_x, (grad_weight if need_w else None), (grad_bias if need_b else None), None, None, None, None, None


class PeerSyncBatchNorm(torch.nn.modules.batchnorm._BatchNorm):
    """Drop-in for torch.nn.SyncBatchNorm (same constructor arguments minus `process_group`, same state_dict keys)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True, comm=None,
                 device=None, dtype=None):
        super().__init__(num_features, eps, momentum, affine, track_running_stats, device=device, dtype=dtype)
        self.comm = comm

    def _check_input_dim(self, x):
        if x.dim() < 2:
            raise ValueError(f'expected at least 2D input (got {x.dim()}D input)')

    def forward(self, x):
        self._check_input_dim(x)
        if self.momentum is None:
            factor = 0.0
        else:
            factor = self.momentum
        if self.training and self.track_running_stats and self.num_batches_tracked is not None:
            self.num_batches_tracked.add_(1)
            if self.momentum is None:  # cumulative moving average
                factor = 1.0 / float(self.num_batches_tracked)
        use_batch_stats = self.training or (self.running_mean is None and self.running_var is None)
        running_mean = self.running_mean if (not self.training or self.track_running_stats) else None
        running_var = self.running_var if (not self.training or self.track_running_stats) else None
        synced = use_batch_stats and self.training and self.comm is not None and self.comm.world > 1
        if not synced:
            return torch.nn.functional.batch_norm(x, running_mean, running_var, self.weight, self.bias, use_batch_stats,
                                                  factor, self.eps)
        if not x.is_cuda:
            raise ValueError('PeerSyncBatchNorm expects input tensors on a CUDA device (dmlcloud_b200 has no CPU path)')
        if x.numel() == 0:
            # torch.nn.SyncBatchNorm filters the rows of ranks with an empty batch out of the gathered statistics with a
            # boolean mask, i.e. a host synchronisation per layer and step (and not at all while capturing); this layer
            # never synchronises, so it refuses the case instead of merging a zero-count row (0 * inf) into the statistics
            raise ValueError('PeerSyncBatchNorm: empty per-rank batch (every rank must contribute at least one sample)')
        return _PeerSyncBN.apply(x, self.weight, self.bias, running_mean, running_var, self.eps, factor, self.comm)


def convert(module, comm):
    """Replace every BatchNorm layer of `module` by a PeerSyncBatchNorm bound to `comm` (parameters, buffers and training
    mode carried over) — the counterpart of torch.nn.SyncBatchNorm.convert_sync_batchnorm used at reference pipeline.py:71."""
    out = module
    if isinstance(module, torch.nn.modules.batchnorm._BatchNorm) and not isinstance(module, PeerSyncBatchNorm):
        out = PeerSyncBatchNorm(module.num_features, module.eps, module.momentum, module.affine,
                                module.track_running_stats, comm=comm)
        if module.affine:
            with torch.no_grad():
                out.weight = module.weight
                out.bias = module.bias
        out.running_mean = module.running_mean
        out.running_var = module.running_var
        out.num_batches_tracked = module.num_batches_tracked
        out.training = module.training
        if hasattr(module, 'qconfig'):
            out.qconfig = module.qconfig
    for name, child in module.named_children():
        out.add_module(name, convert(child, comm))
    return out
