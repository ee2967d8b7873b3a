"""Whole-step CUDA graph for TrainValStage.

The MNIST-CNN step is ~60 kernel launches of a few microseconds each: eager, it is bounded by Python / launch latency,
not by the GPU (SURVEY §3.3 "hot spots").  `GraphedTrainStep` captures one training step

    flat_grad.zero_()  ->  stage.train_step(batch)  [forward, user metrics folded into the slab]
    ->  loss.backward()  ->  gradient sync on the flat bucket  ->  clip (optional)  ->  optimizer.step()
    ->  the stage's per-step metric folds (loss, batch counters)

into ONE cudaGraph and replays it per batch.  Differences from the eager path, all on the gradient side:

  * every parameter's .grad is a VIEW into one flat fp32 bucket (what DDP calls gradient_as_bucket_view), so there is
    no per-parameter copy in or out of a bucket at all;
  * the DDP Reducer is bypassed (`no_sync()`): the flat bucket is synchronised by exactly one libdmlb launch —
    `dmlb_comm_allreduce` (fused scale + cast + NVLink peer all-reduce + write-back, graph-capturable because its
    sequence counter lives in device memory) for W > 1, or the K1/K2 cast round-trip for W == 1 — so the numerics are
    the eager path's (same kernels, same rank-ordered sum);
  * metric folds with python immediates (the int64 batch counters) are baked into the graph — every replay adds 1.

Requirements: static batch shapes; optimizers constructed with `capturable=True` (torch's rule for graph capture);
`step()` must not synchronise with the host (no .item(), no printing of tensors).
"""
import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel

from . import _native as N
from .gradsync import WIRES


class FlatGradBucket:
    """One flat fp32 buffer holding every trainable parameter's gradient; p.grad are views into it."""

    def __init__(self, params, device):
        self.params = [p for p in params if p.requires_grad]
        for p in self.params:
            if p.dtype != torch.float32:
                raise RuntimeError('FlatGradBucket expects fp32 parameters (bf16 autocast keeps fp32 master weights)')
        sizes = [((p.numel() + 3) // 4) * 4 for p in self.params]  # 16-byte aligned slots -> vector path everywhere
        self.total = sum(sizes)
        self.flat = torch.zeros(self.total, dtype=torch.float32, device=device)
        off = 0
        for p, size in zip(self.params, sizes):
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += size

    def attached(self):
        """True while every p.grad still aliases the flat buffer (optimizer.zero_grad(set_to_none=True) would undo it)."""
        base = self.flat.untyped_storage().data_ptr()
        return all(p.grad is not None and p.grad.untyped_storage().data_ptr() == base for p in self.params)


class GraphedTrainStep:
    def __init__(self, stage, example_batch):
        self.stage = stage
        pipeline = stage.pipeline
        self.device = pipeline.device
        if self.device.type != 'cuda':
            raise RuntimeError('cuda_graph mode needs a CUDA device')
        self.lib = N.cuda_lib(self.device.index)
        self.world = dist.get_world_size()
        self.models = list(pipeline.models.values())
        self.ddp_models = [m for m in self.models if isinstance(m, DistributedDataParallel)]
        params, seen = [], set()
        for opt in stage.optimizers():
            for group in opt.param_groups:
                if not group.get('capturable', False):
                    raise RuntimeError('cuda_graph mode: construct the optimizer with capturable=True '
                                       '(e.g. torch.optim.Adam(params, lr=..., capturable=True))')
                for p in group['params']:
                    if id(p) not in seen:
                        seen.add(id(p))
                        params.append(p)
        self.bucket = FlatGradBucket(params, self.device)
        sync = next(iter(pipeline.grad_syncs.values()), None)
        self.wire = sync.wire if sync is not None else pipeline.grad_wire
        self.comm = sync.comm if sync is not None else None
        self.needs_sync = bool(self.ddp_models)
        if self.needs_sync and self.world > 1 and (self.comm is None or
                                                   not self.comm.fits(self._wire_bytes(self.bucket.total))):
            raise RuntimeError('cuda_graph mode needs the peer-memory communicator (grad_route "auto"/"peer") and a '
                               'gradient set that fits grad_arena_bytes')
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=self.device)
        self.static = tuple(torch.empty_like(t, device=self.device) if isinstance(t, torch.Tensor) else t
                            for t in example_batch)
        self.graph = None
        self.loss = None
        self.replays = 0
        self.kernels_in_graph = 0

    def _wire_bytes(self, n):
        return ((n + 7) // 8) * 16 if self.wire == 'bf16' else ((n + 3) // 4) * 16

    # ---- the captured region ---------------------------------------------------------------------------------------
    def _sync_gradients(self):
        flat, n = self.bucket.flat, self.bucket.total
        st = N.stream_ptr()
        clip = bool(self.stage.gradient_clip())
        sumsq_ptr = self.sumsq.data_ptr() if clip else None
        if not self.needs_sync:
            if clip:
                N.check(self.lib.dmlb_bucket_sumsq_f32(flat.data_ptr(), n, sumsq_ptr, st), 'sumsq')
        elif self.world > 1:
            N.check(self.lib.dmlb_comm_allreduce(self.comm.handle, flat.data_ptr(), n, WIRES[self.wire],
                                                 1.0 / self.world, sumsq_ptr, 0, st), 'comm_allreduce')
        elif self.wire == 'bf16':
            N.check(self.lib.dmlb_bucket_round_bf16_f32(flat.data_ptr(), n, 1.0, sumsq_ptr, st), 'round_bf16')
        else:
            N.check(self.lib.dmlb_bucket_scale_f32(flat.data_ptr(), n, 1.0, st), 'scale')
            if clip:
                N.check(self.lib.dmlb_bucket_sumsq_f32(flat.data_ptr(), n, sumsq_ptr, st), 'sumsq')
        if clip:  # one param group == the whole bucket here; the fused sum of squares came for free
            N.check(self.lib.dmlb_bucket_clip_f32(flat.data_ptr(), n, self.sumsq.data_ptr(),
                                                  float(self.stage.gradient_clip()), st), 'clip')

    def _one_step(self):
        stage = self.stage
        self.bucket.flat.zero_()
        self.sumsq.zero_()
        ctxs = [m.no_sync() for m in self.ddp_models]  # the Reducer stays out of it: we synchronise the flat bucket
        for c in ctxs:
            c.__enter__()
        try:
            loss = stage.train_step(self.static)
            loss.backward()
        finally:
            for c in reversed(ctxs):
                c.__exit__(None, None, None)
        self._sync_gradients()
        for opt in stage.optimizers():
            opt.step()
        stage.track_reduce(stage.loss_metric_name(), loss)
        stage._count_batch('train')
        stage.tracker._slab_or_create().flush()  # immediates must be launched INSIDE the capture to be replayed
        return loss

    def capture(self, batch):
        """Capture the step on `batch` (its values are consumed: this is a real training step)."""
        self._load(batch)
        if not self.bucket.attached():
            raise RuntimeError('cuda_graph mode: parameter .grad no longer alias the flat bucket')
        stream = torch.cuda.current_stream(self.device)
        if stream == torch.cuda.default_stream(self.device):
            raise RuntimeError('cuda_graph mode must not run on the legacy default stream (TrainingPipeline.run() puts '
                               'the stages on its compute stream; do the same when driving a stage by hand)')
        # host scalars queued by the last eager step must be launched NOW: flushed inside the capture they would be
        # baked into the graph and re-added by every replay
        self.stage.tracker._slab_or_create().flush()
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        # capture on the very stream the warm-up steps ran on: autograd's AccumulateGrad nodes (stashed by DDP at
        # construction) then already live on the capturing stream and no cross-stream edge enters the graph
        before = N.launch_count()
        with torch.cuda.graph(self.graph, stream=stream):
            self.loss = self._one_step()
        self.kernels_in_graph = N.launch_count() - before  # libdmlb kernels every replay re-runs
        self.graph.replay()  # capture only records: run the step once for real
        self.replays = 1
        return self.loss

    def time_gradient_sync(self, reps=20, per_graph=20):
        """Device time (us) of ONE gradient-sync launch on the flat bucket: `per_graph` of them are captured back to
        back into a throw-away CUDA graph (so host launch latency does not enter) and the replay is timed with CUDA
        events on the launching stream.  Collective: every rank must call it."""
        stream = torch.cuda.current_stream(self.device)
        side = stream if stream != torch.cuda.default_stream(self.device) else torch.cuda.Stream(device=self.device)
        side.wait_stream(stream)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            self._sync_gradients()
            torch.cuda.synchronize(self.device)
            with torch.cuda.graph(g, stream=side):
                for _ in range(per_graph):
                    self._sync_gradients()
            times = []
            for _ in range(reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                g.replay()
                b.record()
                times.append((a, b))
            torch.cuda.synchronize(self.device)
        stream.wait_stream(side)
        return [a.elapsed_time(b) * 1e3 / per_graph for a, b in times]

    def _load(self, batch):
        for dst, src in zip(self.static, batch):
            if isinstance(dst, torch.Tensor):
                dst.copy_(src, non_blocking=True)

    def __call__(self, batch):
        self._load(batch)
        self.graph.replay()
        self.replays += 1
        return self.loss
