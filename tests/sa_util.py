"""Seeded synthetic sparse-attack problem shared by the golden generator and the tests (SURVEY.md §8d config 4 shape:
3x32x32 image, 8x8 grid of 4x4 segments instead of SLIC, eps = 0.1 randn, random-init CifarNet in eval mode)."""
import numpy as np
import torch
from torch import nn


class CifarNet(nn.Module):
    """Same layers, creation order and forward as the attacked model of the reference (SparseAttack/model.py:3-37)."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, kernel_size=3)
        self.conv2 = nn.Conv2d(64, 64, kernel_size=3)
        self.conv3 = nn.Conv2d(64, 128, kernel_size=3)
        self.conv4 = nn.Conv2d(128, 128, kernel_size=3)
        self.pool = nn.MaxPool2d(2, 2)
        self.relu = nn.ReLU(inplace=True)
        self.fc1 = nn.Linear(3200, 256)
        self.dropout = nn.Dropout(0.5)
        self.fc2 = nn.Linear(256, 256)
        self.fc3 = nn.Linear(256, 10)

    def forward(self, x):
        x = self.pool(self.relu(self.conv2(self.relu(self.conv1(x)))))
        x = self.pool(self.relu(self.conv4(self.relu(self.conv3(x)))))
        x = self.relu(self.fc1(x.contiguous().view(-1, 3200)))
        x = self.relu(self.fc2(self.dropout(x)))
        return self.fc3(x)


def grid_segments(block=4, size=32):
    """(size, size) int map of an (size/block)^2 grid of square segments."""
    r = np.arange(size) // block
    return (r[:, None] * (size // block) + r[None, :]).astype(np.int64)


def make_problem(seed=3, n_images=1, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    model = CifarNet().eval()
    images = torch.rand(n_images, 3, 32, 32, generator=g)
    with torch.no_grad():
        target = (model(images - 0.5).argmax(1) + 1) % 10
    eps = 0.1 * torch.randn(n_images, 3, 32, 32, generator=g)
    G0 = torch.ones(n_images, 3, 32, 32)
    seg = grid_segments()
    nseg = int(seg.max()) + 1
    B = torch.zeros(nseg, 3, 32, 32)
    for s in range(nseg):
        B[s, :, torch.from_numpy(seg == s)] = 1
    nw = torch.ones(n_images, 3, 32, 32)
    seg_id = torch.from_numpy(np.broadcast_to(seg, (3, 32, 32)).copy()).reshape(-1).to(torch.int32)
    model = model.to(device)
    return model, images.to(device), target.to(device), eps.to(device), G0.to(device), B.to(device), nw.to(device), seg_id.to(device)


OUTER_CFG = dict(maxIter_e=30, maxIter_g=30, maxIter_mm=1, init_lambda1=0.1, k=1500)


def make_outer_problem(seed=3, n_images=1, device="cpu"):
    """Problem for the outer loop (lambda1 search): a seeded LINEAR 10-class classifier, sensitive enough that
    the six-round search of `train_adptive` takes the x10 branch, succeeds, and bisects within 30 + 30 + 15 iterations per round
    (a random-init CifarNet never flips its label at these iteration counts)."""
    _, images, _, _, _, B, nw, seg_id = make_problem(seed=seed, n_images=n_images)
    g = torch.Generator().manual_seed(seed + 2)
    model = nn.Sequential(nn.Flatten(), nn.Linear(3072, 10))
    with torch.no_grad():
        model[1].weight.copy_(torch.randn(10, 3072, generator=g) * 0.05)
        model[1].bias.zero_()
        target = (model(images - 0.5).argmax(1) + 1) % 10
    model = model.eval().to(device)
    return model, images.to(device), target.to(device), B.to(device), nw.to(device), seg_id.to(device)
