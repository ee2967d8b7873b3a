"""Barebone MNIST with dmlcloud_b200 — a raw `Stage` that owns its loops (counterpart of the reference's
examples/barebone_mnist.py): no TrainValStage, no register_model, hence NO DistributedDataParallel and no gradient
exchange — the model is only moved to the pipeline's device and every rank trains its own replica.  What the stage uses
from the framework is the epoch driver, the progress table and `track_reduce`: the per-batch loss / accuracy go into the
device-resident metric slab (no D2H copy, no sync per batch) and cross the ranks once per epoch in one kernel.

No network here, so the images are synthetic uint8 digits-like data (same shape and dtype as MNIST).

    python examples/barebone_mnist.py                                        # one GPU
    torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node 2 examples/barebone_mnist.py
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch
from torch import nn

from dmlcloud_b200.pipeline import TrainingPipeline
from dmlcloud_b200.stage import Stage
from dmlcloud_b200.util.data import DeviceShardedDataset
from dmlcloud_b200.util.distributed import init_process_group_auto

EPOCHS = int(sys.argv[1]) if len(sys.argv) > 1 else 3


def fake_mnist(count, seed):
    g = torch.Generator().manual_seed(seed)
    labels = torch.randint(0, 10, (count,), generator=g)
    pixels = torch.randint(0, 256, (count, 1, 28, 28), generator=g, dtype=torch.uint8)
    pixels[:, 0, :10, :] = (labels * 25).to(torch.uint8)[:, None, None]  # a learnable signal
    return pixels, labels


class BareboneStage(Stage):
    def pre_stage(self):
        device = self.pipeline.device
        # whole dataset resident in HBM; per-epoch shard indices are the reference's shard_indices, bit for bit
        self.loaders = {
            'train': DeviceShardedDataset(*fake_mnist(8192, 1), batch_size=32, device=device, shuffle=True, seed=0),
            'val': DeviceShardedDataset(*fake_mnist(2048, 2), batch_size=32, device=device, shuffle=False),
        }
        layers = [nn.Conv2d(1, 16, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2), nn.Conv2d(16, 16, 3, padding=1), nn.ReLU(),
                  nn.MaxPool2d(2), nn.Flatten(), nn.Linear(784, 10)]
        self.net = nn.Sequential(*layers).to(device)
        self.opt = torch.optim.Adam(self.net.parameters(), lr=1e-3)
        self.criterion = nn.CrossEntropyLoss()

    def _one_pass(self, split):
        training = split == 'train'
        self.net.train(training)
        self.metric_prefix = split
        loader = self.loaders[split]
        loader.set_epoch(self.current_epoch)
        with torch.set_grad_enabled(training):
            for images, labels in loader:
                logits = self.net(images)
                loss = self.criterion(logits, labels)
                if training:
                    self.opt.zero_grad()
                    loss.backward()
                    self.opt.step()
                self.track_reduce('loss', loss)
                self.track_reduce('accuracy', (logits.argmax(1) == labels).float().mean())

    def run_epoch(self):
        self._one_pass('train')
        self._one_pass('val')

    def table_columns(self):
        extra = [{'name': f'[{split.capitalize()}] {label}', 'metric': f'{split}/{metric}'}
                 for metric, label in (('loss', 'Loss'), ('accuracy', 'Acc.')) for split in ('train', 'val')]
        base = super().table_columns()
        return base[:1] + extra + base[1:]


def main():
    init_process_group_auto()
    pipeline = TrainingPipeline(name='barebone-mnist')
    pipeline.append_stage(BareboneStage(), max_epochs=EPOCHS)
    pipeline.run()
    return pipeline


if __name__ == '__main__':
    main()
