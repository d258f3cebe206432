"""Host-side mirrors of the scoring model (`/root/reference/src/Models/DockingModels.py`) and of the plain-CNN
representation (`/root/reference/src/Models/ProteinRepresentationModels.py:78-129`).

Everything here is ordinary PyTorch: it runs once per protein (representation) or exists for API parity
(`GlobalDockingModel.forward` returning the full score volume).  The search itself is `search.DockingSearch`.
State-dict keys match the reference classes so `.th` checkpoints written by `GlobalDockingModel.save`
(DockingModels.py:55-57) load unchanged.
"""
from __future__ import annotations

import os
from typing import List, Sequence

import torch
from torch import nn

from .volume import VolumeConvolution, VolumeRotation


def _xavier(m):
    if isinstance(m, (nn.Conv3d, nn.Linear)):
        nn.init.xavier_uniform_(m.weight)


class SimpleFilter(nn.Module):
    """Per-translation MLP: Linear(sum(inputs), sum(inputs)//2) -> ReLU -> Linear(., 1)  (DockingModels.py:23-37)."""

    def __init__(self, inputs_sizes: Sequence[int]):
        super().__init__()
        n_in = int(sum(inputs_sizes))
        self.fc_input_size = n_in
        self.fc = nn.Sequential(nn.Linear(n_in, n_in // 2, bias=True), nn.ReLU(), nn.Linear(n_in // 2, 1, bias=True))
        self.fc.apply(_xavier)

    def forward(self, input):
        return self.fc(input)


class E3MultiResRepr4x4(nn.Module):
    """Two-resolution plain CNN: 5-3-3-3-5 convolutions at full resolution, then max-pool(5, stride 2) and
    5-3-3-3 at half resolution; outputs [2m @ L, 4m @ L/2] channels (ProteinRepresentationModels.py:78-129)."""

    def __init__(self, num_input_channels=11, multiplier=16):
        super().__init__()
        c0, c1 = 2 * multiplier, 4 * multiplier
        self.num_outputs_res0, self.num_outputs_res1 = c0, c1

        def stack(cin, cout, kernels):
            layers, c = [], cin
            for i, k in enumerate(kernels):
                layers.append(nn.Conv3d(c, cout, kernel_size=k, padding=k // 2, bias=False))
                if i + 1 < len(kernels):
                    layers.append(nn.ReLU())
                c = cout
            return layers

        self.conv1 = nn.Sequential(*stack(num_input_channels, c0, (5, 3, 3, 3, 5)))
        self.conv2 = nn.Sequential(nn.MaxPool3d(kernel_size=5, stride=2, padding=2), *stack(c0, c1, (5, 3, 3, 3)))

    def get_num_outputs(self):
        return [self.num_outputs_res0, self.num_outputs_res1]

    def forward(self, volume) -> List[torch.Tensor]:
        v0 = self.conv1(volume)
        return [v0, self.conv2(v0)]


class GlobalDockingModel(nn.Module):
    """Container + full-volume scorer with the reference's call surface (DockingModels.py:39-84).

    `forward(receptor_volumes, ligand_volumes)` -> [B,2L,2L,2L]: per-level `VolumeConvolution` (our CUDA
    operator), nearest up-sampling, channel concat, per-voxel filter.  It materialises every intermediate like
    the reference does; `DockingSearch` is the fused path `Docker.dockSE3` uses."""

    def __init__(self, representation, filter, threshold_clash=300, normalize=False, rotate_ligand=False,
                 exclude_clashes=True):
        super().__init__()
        self.threshold_clash = threshold_clash
        self.representation = representation
        self.filter = filter
        self.convolve = VolumeConvolution(clip=5.0)
        self.vol_rotate = VolumeRotation()
        self.rotate_ligand, self.normalize, self.exclude = rotate_ligand, normalize, exclude_clashes

    def save(self, directory, epoch, model_name="DPD_Model"):
        torch.save(self.representation.state_dict(), os.path.join(directory, "%s_repr_epoch%d.th" % (model_name, epoch)))
        torch.save(self.filter.state_dict(), os.path.join(directory, "%s_filter_epoch%d.th" % (model_name, epoch)))

    def load(self, directory, epoch, model_name="DPD_Model"):
        self.representation.load_state_dict(torch.load(os.path.join(directory, "%s_repr_epoch%d.th" % (model_name, epoch))))
        self.filter.load_state_dict(torch.load(os.path.join(directory, "%s_filter_epoch%d.th" % (model_name, epoch))))

    def forward(self, receptor_volumes, ligand_volumes):
        B = receptor_volumes[0].size(0)
        N = 2 * receptor_volumes[0].size(2)
        levels = []
        for rec, lig in zip(receptor_volumes, ligand_volumes):
            corr = self.convolve(rec, lig)
            if corr.size(2) < N:
                corr = nn.functional.interpolate(corr, size=(N, N, N))
            levels.append(corr)
        V = torch.cat(levels, dim=1).permute(0, 2, 3, 4, 1).reshape(B * N * N * N, -1)
        return self.filter(V).reshape(B, N, N, N)


def select_model(group: str = "E3", model: str = "E3MultiResRepr4x4", filter: str = "SimpleFilter", threshold_clash=300):
    """The model table of `local_train.select_model` / `local_test.py` (local_train.py:20-47): equivariance group + model name
    -> `GlobalDockingModel(representation, SimpleFilter, threshold_clash)` with the reference's multiplier of 8."""
    from .se3_scalar import SE3MultiResReprScalar
    if group == "E3":
        if model != "E3MultiResRepr4x4":
            raise Exception("Unknown model name", model)
        representation = E3MultiResRepr4x4(multiplier=8)
    elif group == "SE3":
        if model != "SE3MultiResReprScalar":
            raise Exception("Unknown model name", model)
        representation = SE3MultiResReprScalar(multiplier=8)
    else:
        raise Exception("Unknown equivariance group", group)
    if filter != "SimpleFilter":
        raise Exception("Unknown filter name", filter)
    return GlobalDockingModel(representation, SimpleFilter(representation.get_num_outputs()), threshold_clash=threshold_clash)
