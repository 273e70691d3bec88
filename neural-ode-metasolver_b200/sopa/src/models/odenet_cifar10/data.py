"""CIFAR-10 input statistics and the on-GPU form of the training augmentation of the reference data pipeline
(sopa/src/models/odenet_cifar10/data.py:40-57: RandomCrop(32, padding=4) + RandomHorizontalFlip + ToTensor +
Normalize(CIFAR_MEAN, CIFAR_STD)).  The dataset readers themselves (torchvision CIFAR10, DataLoader workers) are out of
scope (drivers, not the path); what the hot path's callers need on the device is the normalisation constants
and a batched augmentation that does not bounce through the host."""
import torch

CIFAR_MEAN = (0.4914, 0.4822, 0.4465)      # data.py:45
CIFAR_STD = (0.2023, 0.1994, 0.2010)


def normalize(img01):
    """(img - mean) / std for a (B,3,H,W) batch in [0,1] (transforms.Normalize, data.py:45)."""
    mean = torch.tensor(CIFAR_MEAN, dtype=img01.dtype, device=img01.device).view(1, 3, 1, 1)
    std = torch.tensor(CIFAR_STD, dtype=img01.dtype, device=img01.device).view(1, 3, 1, 1)
    return (img01 - mean) / std


def augment_batch(img01, generator=None, padding=4):
    """RandomCrop(32, padding=4) + RandomHorizontalFlip() (data.py:41-43) for a whole (B,3,H,W) batch ON ITS DEVICE:
    one zero-padded copy, per-sample crop offsets / flip bits drawn from `generator` (a torch.Generator on the same
    device, or None), one gather.  Returns the augmented batch in [0,1] (normalise afterwards)."""
    B, C, H, W = img01.shape
    dev = img01.device
    dx = torch.randint(0, 2 * padding + 1, (B,), device=dev, generator=generator)
    dy = torch.randint(0, 2 * padding + 1, (B,), device=dev, generator=generator)
    flip = torch.rand(B, device=dev, generator=generator) < 0.5
    padded = torch.nn.functional.pad(img01, (padding, padding, padding, padding))
    ar_h = torch.arange(H, device=dev).view(1, H, 1)
    ar_w = torch.arange(W, device=dev).view(1, 1, W)
    rows = (dy.view(B, 1, 1) + ar_h).expand(B, H, W)
    cols_fwd = dx.view(B, 1, 1) + ar_w
    cols = torch.where(flip.view(B, 1, 1), dx.view(B, 1, 1) + (W - 1 - ar_w), cols_fwd).expand(B, H, W)
    idx = (rows * (W + 2 * padding) + cols).view(B, 1, H * W).expand(B, C, H * W)
    return padded.reshape(B, C, -1).gather(2, idx).view(B, C, H, W)
