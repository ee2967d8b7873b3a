"""Launches the fused step exchange (dmlb_comm_allreduce with a dmlb_step_metrics descriptor) on a local (W = 1) communicator
through the C ABI: the MNIST flat bucket (10,332 elements: launch-latency bound) and the ResNet-18 flat bucket
(11,689,512 elements: 8 B/element in place, HBM/L2 bound).  Target of the ncu captures of round 2:

    ncu --set full --clock-control none --import-source on -k regex:allreduce_oneshot -c 6 \
        -o gpurun_out/prof_step_exchange python profiles/run_step_exchange.py
"""
import ctypes
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from dmlcloud_b200 import _native as N  # noqa: E402
from dmlcloud_b200.gradsync import PeerComm  # noqa: E402
from dmlcloud_b200.metrics import DeviceSlab, HostFeed, StepRing, _layout_hash  # noqa: E402

dev = torch.device('cuda', 0)
lib = N.cuda_lib(0)
comm = PeerComm(dev, max_message_bytes=1 << 16)
slab = DeviceSlab(dev)
cells = [slab.alloc(1, d) for d in (0 | 8, 0 | 8, 1 | 4 | 8, 1 | 4, 0 | 8)]  # loss, accuracy, total, worker (local), step time
ring, feed = StepRing(lib, slab.capacity), HostFeed(lib)
feed.assign({cells[4]: (N.MEAN, False)})
counter = torch.zeros(1, dtype=torch.int64, device=dev)
loss, acc = torch.rand((), device=dev), torch.rand((), device=dev)
st = N.stream_ptr()
for n in (10_332, 11_689_512):
    bucket = torch.randn(n, device=dev)
    for rep in range(3):
        feed.put(cells[4], 0.25)
        feed.commit(int(counter.item()))
        m = N.StepMetrics()
        m.acc, m.cnt, m.desc = slab.acc.data_ptr(), slab.cnt.data_ptr(), slab.desc.data_ptr()
        m.counter, m.out_ring, m.feed = counter.data_ptr(), ring.device_ptr, feed.device_ptr
        m.layout_hash, m.n_cells, m.capacity = _layout_hash('profile'), slab.n_cells, slab.capacity
        m.ring_slots, m.feed_slots = StepRing.SLOTS, HostFeed.SLOTS
        folds = [N.FoldEntry(loss.data_ptr(), 0, N.F32, cells[0], 1, 1, 1, 0), N.FoldEntry(acc.data_ptr(), 0, N.F32, cells[1], 1, 1, 1, 0),
                 N.FoldEntry(None, 1, N.F64, cells[2], 1, 1, 1, 0), N.FoldEntry(None, 1, N.F64, cells[3], 1, 1, 1, 0),
                 N.FoldEntry(None, 0, N.SRC_FEED, cells[4], 1, 0, 1, 0)]
        m.n_folds = len(folds)
        for i, e in enumerate(folds):
            m.folds[i] = e
        ranges = [(cells[0], cells[2] + 1), (cells[4], cells[4] + 1), (cells[3], cells[3] + 1)]
        m.n_ranges, m.n_global_ranges = 3, 2
        for i, (b, e) in enumerate(ranges):
            m.ranges[i] = N.Range(b, e)
        N.check(lib.dmlb_comm_allreduce(comm.handle, bucket.data_ptr(), n, N.WIRE_BF16, 1.0, None, 0, ctypes.byref(m), st))
        torch.cuda.synchronize()
print('done', N.launch_count(), 'stamp', ring.latest())
