"""Whole-step CUDA graph for TrainValStage with the fused step exchange.

The MNIST-CNN step is ~60 kernel launches of a few microseconds each: eager, it is bounded by Python / launch latency,
not by the GPU (SURVEY §3.3 "hot spots").  `GraphedTrainStep` captures one training step

    flat_grad.zero_()  ->  stage.train_step(batch)  [forward; user metrics are QUEUED, not launched]
    ->  loss.backward()  ->  ONE libdmlb launch = gradient all-reduce on the flat bucket  +  the step's metric folds
        +  the cross-rank exchange of the running metric values (fused step exchange, csrc/peer_comm.cu)
    ->  optimizer.step()  [clip coefficient fused into the FlatAdam / FlatSGD kernel]

into ONE cudaGraph and replays it per batch.  What is different from the eager loop (stage.py train_epoch):

  * every parameter's .grad is a VIEW into one flat fp32 bucket (what DDP calls gradient_as_bucket_view), so there is
    no per-parameter copy in or out of a bucket at all; the DDP Reducer is bypassed (`no_sync()`);
  * the per-step metric traffic of the reference — 4x track_reduce + the user's (stage.py:305-314) and, in the
    per-step operating point of BASELINE configs 2/3, a cross-rank reduction of all of them — costs NO launch and NO
    barrier of its own: one extra CTA of the all-reduce kernel folds the values, exchanges 16-byte records under the
    gradients' flag barrier and writes the results into a ring in mapped host memory (`stage.live_metrics`);
  * host scalars tracked between replays (misc/step_time_ms, stage.py:314) travel INTO the graph through a second ring
    in mapped host memory (metrics.HostFeed), one slot per replay — no launch, no copy;
  * learning rates live in device memory (optim.FlatAdam / FlatSGD), so `scheduler.step()` (stage.py:316-318) takes
    effect on the next replay; a torch optimizer with a python-float lr would have it baked in, which is refused.

Requirements: static batch shapes; FlatAdam / FlatSGD, or torch optimizers constructed with `capturable=True` and no
scheduler; `step()` must not synchronise with the host (no .item(), no printing of tensors).
"""
import ctypes

import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel

from . import _native as N
from .gradsync import WIRES, PeerComm
from .metrics import HostFeed, StepRing, _RingResult


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
        params, seen, groups = [], set(), 0
        for opt in stage.optimizers():
            device_lr = getattr(opt, 'device_lr', False)
            for group in opt.param_groups:
                groups += 1
                if not device_lr and not group.get('capturable', False):
                    raise RuntimeError('cuda_graph mode: use dmlcloud_b200.optim.FlatAdam / FlatSGD, or construct the '
                                       'torch optimizer with capturable=True')
                if not device_lr and not isinstance(group.get('lr'), torch.Tensor) and pipeline.schedulers:
                    raise RuntimeError('cuda_graph mode: a python-float learning rate is baked into the captured graph, '
                                       'so the registered scheduler would be silently ignored; use FlatAdam / FlatSGD '
                                       '(device-resident lr) or a tensor lr')
                for p in group['params']:
                    if id(p) not in seen:
                        seen.add(id(p))
                        params.append(p)
        self.clip = float(stage.gradient_clip() or 0.0)
        if self.clip and groups != 1:
            # the reference clips per param group (stage.py:276-279); the fused sum of squares covers the whole flat bucket
            raise RuntimeError('cuda_graph mode with gradient_clip() supports exactly one optimizer param group')
        self.bucket = FlatGradBucket(params, self.device)
        sync = next(iter(pipeline.grad_syncs.values()), None)
        self.wire = sync.wire if sync is not None else pipeline.grad_wire
        self.algo = sync.algo if sync is not None else 0
        self.needs_sync = bool(self.ddp_models)
        self._own_comm = None
        comm = sync.comm if (sync is not None and self.needs_sync) else None
        if comm is None and self.world > 1 and not self.needs_sync:
            comm = pipeline.metric_comm  # no gradients to exchange: the metric records ride on the metric communicator
        if comm is None and self.world == 1:
            comm = self._own_comm = PeerComm(self.device, max_message_bytes=1 << 16)  # local: no peers, no mapping
        if comm is None or (self.needs_sync and not comm.fits(self._wire_bytes(self.bucket.total))):
            raise RuntimeError('cuda_graph mode needs the peer-memory communicator (grad_route "auto"/"peer") and a '
                               'gradient set that fits grad_arena_bytes')
        self.comm = comm
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=self.device)
        self.static = tuple(torch.empty_like(t, device=self.device) if isinstance(t, torch.Tensor) else t
                            for t in example_batch)
        self.graph = None
        self.loss = None
        self.replays = 0
        self.kernels_in_graph = 0
        # fused step exchange state (built at capture)
        self.ring = None
        self.feed = None
        self.counter = None
        self.replays_at_ring = 0
        self.live_names = {}
        self.zero_in_optimizer = False
        self._copy_stream = None  # side stream + double-buffered staging for large pinned host batches (see _load)
        self._staging = {}
        self._slab_generation = None
        self._keep = None
        self.step_metrics = None

    def _wire_bytes(self, n):
        return ((n + 7) // 8) * 16 if self.wire == 'bf16' else ((n + 3) // 4) * 16

    # ---- the fused step exchange -------------------------------------------------------------------------------------
    def _describe_metrics(self, entries):
        """dmlb_step_metrics for this step: the queued fold entries + the live selection + the two host rings."""
        tracker = self.stage.tracker
        slab = tracker._slab_or_create()
        by_name, plan = tracker.live_selection()
        if not by_name:
            return None
        glob, loc, layout = plan
        n_glob = sum(e - b for b, e in glob)
        if n_glob > N.STEP_METRIC_MAX_CELLS or len(glob) + len(loc) > N.MAX_RANGES or len(entries) > N.MAX_FOLD_ENTRIES:
            return None  # too large for the piggy-back: the stage falls back to the separate exchange kernel
        m = N.StepMetrics()
        m.acc, m.cnt, m.desc = slab.acc.data_ptr(), slab.cnt.data_ptr(), slab.desc.data_ptr()
        m.counter = self.counter.data_ptr()
        m.out_ring = self.ring.device_ptr
        m.feed = self.feed.device_ptr if self.feed is not None else None
        m.layout_hash = layout
        m.n_cells, m.capacity = slab.n_cells, slab.capacity
        m.ring_slots, m.feed_slots = StepRing.SLOTS, (HostFeed.SLOTS if self.feed is not None else 0)
        m.n_folds = len(entries)
        for i, e in enumerate(entries):
            m.folds[i] = e
        ranges = list(glob) + list(loc)
        m.n_ranges, m.n_global_ranges = len(ranges), len(glob)
        for i, (b, e) in enumerate(ranges):
            m.ranges[i] = N.Range(b, e)
        self.live_names = dict(by_name)
        return m

    def _sync_gradients(self, metrics=None):
        flat, n = self.bucket.flat, self.bucket.total
        st = N.stream_ptr()
        sumsq_ptr = self.sumsq.data_ptr() if self.clip else None
        if self.needs_sync or metrics is not None:
            N.check(self.lib.dmlb_comm_allreduce(self.comm.handle, flat.data_ptr() if self.needs_sync else None,
                                                 n if self.needs_sync else 0, WIRES[self.wire], 1.0 / self.world,
                                                 sumsq_ptr if self.needs_sync else None, self.algo,
                                                 ctypes.byref(metrics) if metrics is not None else None, st),
                    'comm_allreduce')
        if self.clip and not self.needs_sync:
            N.check(self.lib.dmlb_bucket_sumsq_f32(flat.data_ptr(), n, sumsq_ptr, st), 'sumsq')

    def _optimize(self):
        stage = self.stage
        clip = (self.sumsq, self.clip) if self.clip else None
        for opt in stage.optimizers():
            if clip is not None and getattr(opt, 'fused_clip', False):
                opt.step(clip=clip)  # coefficient derived on the device inside the K5 / K6 launch: no extra pass
                clip = None
            else:
                if clip is not None:
                    N.check(self.lib.dmlb_bucket_clip_f32(self.bucket.flat.data_ptr(), self.bucket.total,
                                                          self.sumsq.data_ptr(), self.clip, N.stream_ptr()), 'clip')
                    clip = None
                opt.step()

    def _one_step(self):
        stage = self.stage
        slab = stage.tracker._slab_or_create()
        if not self.zero_in_optimizer:
            self.bucket.flat.zero_()  # (FlatAdam / FlatSGD zero the gradients they consumed inside their own launch)
        if self.clip:
            self.sumsq.zero_()
        slab.batching = True
        try:
            ctxs = [m.no_sync() for m in self.ddp_models]  # the Reducer stays out of it: we synchronise the flat bucket
            for c in ctxs:
                c.__enter__()
            try:
                loss = stage.train_step(self.static)
                loss.backward()
            finally:
                for c in reversed(ctxs):
                    c.__exit__(None, None, None)
            stage.track_reduce(stage.loss_metric_name(), loss)
            stage._count_batch('train')
            entries, keep = slab.take_batch()
        finally:
            slab.batching = False
        if self.feed is not None:
            # python scalars the stage tracks BETWEEN steps (misc/step_time_ms, stage.py:314) get a column of the host feed
            # ring; the ones tracked inside the step (the batch counters) are immediates of this very fold
            inside = {e.cell for e in entries if not e.src}
            cols = {c: kind for c, kind in slab.imm_cells_seen.items() if c not in inside}
            room = N.MAX_FOLD_ENTRIES - len(entries)
            cols = dict(sorted(cols.items())[:max(0, min(N.FEED_WIDTH, room))])
            self.feed.assign(cols)
            for cell, j in self.feed.cols.items():
                entries.append(N.FoldEntry(None, 0, N.SRC_FEED, cell, 1, j, 1, 0))
        metrics = self._describe_metrics(entries) if stage.live_metrics_every else None
        if metrics is None:  # no live exchange wanted (or it does not fit): plain fold launch(es), inside the graph
            if self.feed is not None:
                self.feed.assign({})  # nobody would read the feed ring: python scalars keep their normal route
            real = [e for e in entries if e.src_dtype != N.SRC_FEED]
            for i in range(0, len(real), N.MAX_FOLD_ENTRIES):
                slab._launch_fold(real[i:i + N.MAX_FOLD_ENTRIES])
        self.step_metrics, self._keep = metrics, keep
        self._sync_gradients(metrics)
        self._optimize()
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
        stage = self.stage
        slab = stage.tracker._slab_or_create()
        # host scalars queued by the last eager step must be launched NOW: inside the capture they would be baked into
        # the graph and re-added by every replay
        slab.flush_all()  # (also hands scalars still waiting in a previous capture's feed ring to a normal fold launch)
        slab.feed = None
        for opt in stage.optimizers():
            sync_lr = getattr(opt, 'sync_device_lr', None)
            if sync_lr is not None:
                sync_lr()
        # pinned memory cannot be allocated while a stream is capturing: the host feed ring exists before the capture, its
        # columns are assigned inside it (when the step has shown which python scalars it tracks itself)
        self.feed = HostFeed(self.lib) if stage.live_metrics_every else None
        # result ring (mapped host memory) and exchange counter (device) of the fused step exchange: created OUTSIDE the
        # capture — a tensor made inside it would be re-initialised by every replay
        self.ring = StepRing(self.lib, slab.capacity) if stage.live_metrics_every else None
        self.counter = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.replays_at_ring = self.replays
        torch.cuda.synchronize(self.device)
        # `optimizer.zero_grad()` of the next step (reference stage.py:300) is fused into the K5 / K6 launch when ONE flat
        # optimizer owns every gradient of the bucket: one kernel node and one pass over the gradients fewer per step
        opts = list(stage.optimizers())
        self.zero_in_optimizer = (len(opts) == 1 and getattr(opts[0], 'device_lr', False) and len(opts[0].param_groups) == 1
                                  and opts[0]._flat_grad_base(opts[0].param_groups[0], opts[0]._flat[0]) ==
                                  self.bucket.flat.data_ptr() and opts[0]._flat[0]['total'] == self.bucket.total)
        if self.zero_in_optimizer:
            opts[0].zero_grad_in_step = True
            self.bucket.flat.zero_()  # once, outside the graph: every replay leaves zeros behind
        self.graph = torch.cuda.CUDAGraph()
        # capture on the very stream the warm-up steps ran on: autograd's AccumulateGrad nodes (stashed by DDP at
        # construction) then already live on the capturing stream and no cross-stream edge enters the graph
        before = N.launch_count()
        with torch.cuda.graph(self.graph, stream=stream):
            self.loss = self._one_step()
        self.kernels_in_graph = N.launch_count() - before  # libdmlb kernels every replay re-runs
        self._slab_generation = slab.generation
        slab.feed = self.feed  # from now on python scalars of the feed's cells wait for the next replay
        self._replay()  # capture only records: run the step once for real
        return self.loss

    def _replay(self):
        if self.feed is not None and self.step_metrics is not None:
            count = self.replays_since_ring()  # exchanges issued so far == the slot index the device will use
            if count % 16 == 0 and count - self.counter_host() >= HostFeed.SLOTS // 2:
                # the host is half a ring ahead of the GPU: wait for the exchange that frees the slot about to be written
                self.ring.wait(count - HostFeed.SLOTS // 2 + 1, sync=lambda: torch.cuda.synchronize(self.device))
            self.feed.commit(count)
        self.graph.replay()
        self.replays += 1
        if self.step_metrics is not None:
            k = self.replays_since_ring()
            self.stage.live_metrics = self.stage.tracker.live_view(
                _RingResult(self.ring, k, sync=lambda: torch.cuda.synchronize(self.device)), self.live_names)

    def replays_since_ring(self):
        return self.replays - self.replays_at_ring

    def counter_host(self):
        """Number of step exchanges the GPU has completed, read from the result ring's stamps (no CUDA call)."""
        return self.ring.latest() if self.ring is not None else 0

    def time_gradient_sync(self, reps=20, per_graph=20):
        """Device time (us) of ONE gradient-sync launch on the flat bucket (without the metric CTA): `per_graph` of them
        are captured back to back into a throw-away CUDA graph (so host launch latency does not enter) and the replay is
        timed with CUDA events on the launching stream.  Collective: every rank must call it."""
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

    STAGE_MIN_BYTES = 1 << 20

    def _load(self, batch):
        """Bring `batch` into the captured step's static input buffers.  Large batches that sit in PINNED host memory
        (ResNet-18: 38.5 MB per step) take a detour that hides the PCIe transfer: the H2D copy goes to one of two staging
        buffers on a copy stream — the host issues it while the GPU is still computing the previous step — and the
        compute stream only does a device-to-device copy (microseconds) once the staged data has landed."""
        for i, (dst, src) in enumerate(zip(self.static, batch)):
            if not isinstance(dst, torch.Tensor):
                continue
            staged = (isinstance(src, torch.Tensor) and not src.is_cuda and src.is_pinned()
                      and src.numel() * src.element_size() >= self.STAGE_MIN_BYTES)
            if not staged:
                dst.copy_(src, non_blocking=True)
                continue
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=self.device)
            slot = self._staging.get(i)
            if slot is None:
                slot = self._staging[i] = {'buf': [torch.empty_like(dst), torch.empty_like(dst)], 'next': 0,
                                           'ready': [torch.cuda.Event(), torch.cuda.Event()],
                                           'consumed': [torch.cuda.Event(), torch.cuda.Event()], 'used': [False, False]}
            k = slot['next']
            slot['next'] ^= 1
            compute = torch.cuda.current_stream(self.device)
            if slot['used'][k]:
                self._copy_stream.wait_event(slot['consumed'][k])  # the step that last read this staging buffer has copied it out
            with torch.cuda.stream(self._copy_stream):
                slot['buf'][k].copy_(src, non_blocking=True)
                slot['ready'][k].record(self._copy_stream)
            compute.wait_event(slot['ready'][k])
            dst.copy_(slot['buf'][k], non_blocking=True)
            slot['consumed'][k].record(compute)
            slot['used'][k] = True

    def __call__(self, batch):
        slab = self.stage.tracker._slab_or_create()
        if slab.generation != self._slab_generation:
            # the metric slab was reallocated (it grew): the graph holds stale pointers -> capture again on this batch
            return self.capture(batch)
        for opt in self.stage.optimizers():
            sync_lr = getattr(opt, 'sync_device_lr', None)
            if sync_lr is not None:
                sync_lr()  # a scheduler changed group['lr']: one tiny fill, only when the value actually changed
        self._load(batch)
        self._replay()
        return self.loss

    def detach(self):
        """End of the stage: scalars still waiting for a replay take the normal route; later stages see a plain slab."""
        slab = self.stage.tracker._slab
        if slab is not None and slab.feed is not None and slab.feed is self.feed:
            slab.flush_all()
            slab.feed = None

    def close(self):
        self.detach()
        if self._own_comm is not None:
            self._own_comm.close()
            self._own_comm = None
