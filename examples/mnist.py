"""MNIST CNN with dmlcloud_b200 — the reference's examples/mnist.py, line for line where the API is the same, with the
B200 extras switched on: device-resident sharded dataset, bf16 gradient wire, whole-step CUDA graph, Adam as one
launch on flat buffers, per-step metric exchange.  No network here, so the images are synthetic uint8 (same shape and dtype as MNIST).

    python examples/mnist.py                                        # one GPU
    torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node 8 examples/mnist.py
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import torch
from torch import nn

from dmlcloud_b200.optim import FlatAdam
from dmlcloud_b200.pipeline import TrainingPipeline
from dmlcloud_b200.stage import TrainValStage
from dmlcloud_b200.util.data import DeviceShardedDataset
from dmlcloud_b200.util.distributed import init_process_group_auto


def synthetic_mnist(n, seed):
    g = torch.Generator().manual_seed(seed)
    labels = torch.randint(0, 10, (n,), generator=g)
    images = torch.randint(0, 256, (n, 1, 28, 28), generator=g, dtype=torch.uint8)
    images[:, 0, :10, :] = (labels * 25).to(torch.uint8)[:, None, None]  # make the task learnable
    return images, labels


class MNISTStage(TrainValStage):
    def pre_stage(self):
        train_x, train_y = synthetic_mnist(60000, seed=0)
        val_x, val_y = synthetic_mnist(10000, seed=1)
        # whole dataset lives in HBM; per step one gather + normalise kernel (reference: DataLoader + ToTensor + Normalize)
        self.pipeline.register_dataset('train', DeviceShardedDataset(train_x, train_y, batch_size=32, shuffle=True,
                                                                     device=self.device, drop_last=True))
        self.pipeline.register_dataset('val', DeviceShardedDataset(val_x, val_y, batch_size=32, shuffle=False,
                                                                   device=self.device, drop_last=True))
        model = nn.Sequential(
            nn.Conv2d(1, 16, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2),
            nn.Conv2d(16, 16, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2),
            nn.Flatten(), nn.Linear(784, 10),
        )
        self.pipeline.register_model('cnn', model, grad_wire='bf16')
        # the reference registers torch.optim.Adam(model.parameters(), lr=1e-3); that still works here (pass
        # capturable=True for the captured step).  FlatAdam is the same optimizer as ONE libdmlb launch per step.
        self.pipeline.register_optimizer('adam', FlatAdam(model.parameters(), lr=1e-3))
        self.loss = nn.CrossEntropyLoss()
        self.cuda_graph = True          # capture the whole step after 3 eager steps
        self.live_metrics_every = 50    # running metrics cross the ranks every 50 steps (one fused kernel)

    def step(self, batch) -> torch.Tensor:
        img, target = batch             # already on the device
        with torch.autocast('cuda', dtype=torch.bfloat16):
            output = self.pipeline.models['cnn'](img)
        loss = self.loss(output.float(), target)
        self.track_reduce('accuracy', (output.argmax(1) == target).float().mean())
        return loss

    def table_columns(self):
        columns = super().table_columns()
        columns.insert(-2, {'name': '[Val] Acc.', 'metric': 'val/accuracy'})
        columns.insert(-2, {'name': '[Train] Acc.', 'metric': 'train/accuracy'})
        return columns


def main():
    init_process_group_auto()
    pipeline = TrainingPipeline(name='mnist')
    pipeline.append_stage(MNISTStage(), max_epochs=3)
    pipeline.run()


if __name__ == '__main__':
    main()
