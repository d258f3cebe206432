"""Plain-PyTorch restatement of the scalar-only SE(3) representation network
(`/root/reference/src/Models/ProteinRepresentationModels.py:23-76`, selected by `local_train.py:27-31`).

The reference builds it from `se3cnn.SE3Convolution([(n_in, 0)], [(n_out, 0)], size=5, padding=2, dyn_iso=True, bias=None)`
and `ScalarActivation([(n, relu)], bias=False)`.  With only l = 0 fields on both sides an SE(3)-equivariant kernel is a
radial function: se3cnn expands it in a small basis of Gaussian radial shells and learns one coefficient per
(output channel, input channel, shell):

    K[o, i, x] = sum_beta  weight[o, i, beta] * basis[beta, x]          (se3cnn SE3Kernel.combination)
    out        = conv3d(in, K, padding, stride)                         (se3cnn SE3Convolution.forward)

se3cnn is not vendored with the reference and cannot be installed here (no network): PARITY UNPINNED for the basis
construction (shell radii 0, 1, 2 voxels, sigma = 0.6, unit L2 norm per shell) and for the `dyn_iso` initialisation.
What is pinned is the state-dict layout: se3cnn stores the basis as a *buffer* next to the weight vector, so a real `.th`
checkpoint carries its own basis and `load_state_dict` restores both --

    sequence_res{0,1}.{0,2,4,6}.kernel.weight        [n_out * n_in * n_basis]
    sequence_res{0,1}.{0,2,4,6}.kernel.kernel_0_0    [n_basis, 1, 1, 5, 5, 5]

after which the forward pass depends on nothing constructed in this file.  The network runs once per protein, outside the
rotation loop (BASELINE.json north_star: "stays as plain PyTorch").
"""
from __future__ import annotations

import math
from typing import List

import torch
from torch import nn
import torch.nn.functional as F


def radial_shell_basis(size: int = 5, sigma: float = 0.6) -> torch.Tensor:
    """[size // 2 + 1, 1, 1, size, size, size]: Gaussian shells exp(-((|x| - r) / sigma)^2 / 2) at r = 0 .. size // 2 voxels,
    each scaled to unit L2 norm (the l_in = l_out = 0 block of se3cnn's `cube_basis_kernels_analytical`)."""
    ax = torch.arange(size, dtype=torch.float64) - (size - 1) / 2.0
    X, Y, Z = torch.meshgrid(ax, ax, ax, indexing="ij")
    r = torch.sqrt(X * X + Y * Y + Z * Z)
    shells = []
    for radius in range(size // 2 + 1):
        w = torch.exp(-0.5 * ((r - radius) / sigma) ** 2) / (math.sqrt(2.0 * math.pi) * sigma)
        shells.append(w / w.norm())
    return torch.stack(shells).reshape(len(shells), 1, 1, size, size, size).float()


class ScalarSE3Kernel(nn.Module):
    """`se3cnn.SE3Kernel` restricted to scalar fields: `weight` [n_out * n_in * n_basis] and the basis buffer `kernel_0_0`."""

    def __init__(self, n_in: int, n_out: int, size: int = 5):
        super().__init__()
        self.n_in, self.n_out, self.size = int(n_in), int(n_out), int(size)
        basis = radial_shell_basis(size)
        self.register_buffer("kernel_0_0", basis)
        nb = basis.shape[0]
        # variance-preserving start (stands in for se3cnn's dyn_iso initialisation; overwritten by any checkpoint)
        self.weight = nn.Parameter(torch.randn(self.n_out * self.n_in * nb) / math.sqrt(self.n_in * nb))

    def forward(self) -> torch.Tensor:
        nb = self.kernel_0_0.shape[0]
        w = self.weight.reshape(self.n_out * self.n_in, nb)
        k = w @ self.kernel_0_0.reshape(nb, -1)
        return k.reshape(self.n_out, self.n_in, self.size, self.size, self.size)


class ScalarSE3Convolution(nn.Module):
    """`se3cnn.SE3Convolution([(n_in, 0)], [(n_out, 0)], size, padding=, stride=, bias=None)`."""

    def __init__(self, n_in: int, n_out: int, size: int = 5, padding: int = 2, stride: int = 1):
        super().__init__()
        self.kernel = ScalarSE3Kernel(n_in, n_out, size)
        self.padding, self.stride = int(padding), int(stride)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return F.conv3d(x, self.kernel(), bias=None, padding=self.padding, stride=self.stride)


class ScalarReLU(nn.Module):
    """`se3cnn.non_linearities.ScalarActivation([(n, F.relu)], bias=False)`: no parameters."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return F.relu(x)


class SE3MultiResReprScalar(nn.Module):
    """Two-resolution scalar-field SE(3) network: four radial 5^3 convolutions at full resolution (11 -> 2m channels), then a
    stride-2 radial convolution and three more at half resolution (2m -> 4m); returns [2m @ L, 4m @ L/2]
    (ProteinRepresentationModels.py:23-76; multiplier = 8 in local_train.py:30)."""

    def __init__(self, num_input_channels: int = 11, multiplier: int = 16):
        super().__init__()
        c0, c1 = 2 * multiplier, 4 * multiplier
        self.num_outputs_res0, self.num_outputs_res1 = c0, c1
        self.sequence_res0 = nn.Sequential(
            ScalarSE3Convolution(num_input_channels, c0), ScalarReLU(),
            ScalarSE3Convolution(c0, c0), ScalarReLU(),
            ScalarSE3Convolution(c0, c0), ScalarReLU(),
            ScalarSE3Convolution(c0, c0))
        self.sequence_res1 = nn.Sequential(
            ScalarSE3Convolution(c0, c1, stride=2), ScalarReLU(),
            ScalarSE3Convolution(c1, c1), ScalarReLU(),
            ScalarSE3Convolution(c1, c1), ScalarReLU(),
            ScalarSE3Convolution(c1, c1))

    def get_num_outputs(self) -> List[int]:
        return [self.num_outputs_res0, self.num_outputs_res1]

    def forward(self, volume: torch.Tensor) -> List[torch.Tensor]:
        vol1 = self.sequence_res0(volume)
        vol2 = self.sequence_res1(vol1)
        return [vol1, vol2]
