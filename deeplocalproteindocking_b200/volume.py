"""Drop-ins for the `TorchProteinLibrary.Volume` operators the docking path imports
(`/root/reference/src/Docker/Docker.py:9`, `/root/reference/src/Models/DockingModels.py:13`).

`VolumeRotation` and `VolumeConvolution` call the hand-written CUDA kernels in libdlpd.so through
`torch.autograd.Function`s (forward only: the search runs under `torch.no_grad()`,
`/root/reference/src/local_test.py:67`; asking for a gradient raises).  `TypedCoords2Volume` runs the CUDA projector
(`dlpd_project_typed`) for float32 CUDA inputs -- the receptor / ligand set-up and the per-rotation projection of the
`dockE3` loop (Docker.py:164,204,208) -- and keeps a plain-PyTorch form for CPU tensors (host-side tests only).
"""
from __future__ import annotations

import ctypes

import torch
from torch import nn

from . import _lib
from .search import _ptr, _stream_ptr


def _require_cuda_f32(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.DlpdError("%s must live on a CUDA device: libdlpd has no CPU path" % what)
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32, got %s" % (what, t.dtype))
    return t.contiguous()


class _NoBackward:
    @staticmethod
    def backward(ctx, *grads):
        raise NotImplementedError("libdlpd operators are forward-only (the docking search runs under no_grad)")


class VolumeRotationFunction(_NoBackward, torch.autograd.Function):
    @staticmethod
    def forward(ctx, volume, R, center_mode):
        volume = _require_cuda_f32(volume, "volume")
        R = _require_cuda_f32(R, "R")
        B, C, L = volume.shape[0], volume.shape[1], volume.shape[2]
        if volume.dim() != 5 or volume.shape[3] != L or volume.shape[4] != L or tuple(R.shape) != (B, 3, 3):
            raise ValueError("expected volume [B,C,L,L,L] and R [B,3,3]")
        out = torch.empty_like(volume)
        lib = _lib.get_lib()
        with torch.cuda.device(volume.device):
            _lib.check(lib, lib.dlpd_volume_rotate(_ptr(volume), _ptr(R), _ptr(out), B, C, L, center_mode,
                                                  _stream_ptr(volume.device)))
        return out


class VolumeRotation(nn.Module):
    """`TorchProteinLibrary.Volume.VolumeRotation` (Docker.py:40,218): `forward(volume[B,C,L,L,L], R[B,3,3])`.

    out[b,c,p] = trilinear(volume[b,c], ctr + R_b^T (p - ctr)), zeros outside; ctr = (L-1)/2 ('grid', default: what
    `affine_grid` + `grid_sample` -- the construction TorchProteinLibrary's operator is built on -- does, see
    tests/test_host_logic.py) or L/2 ('box', opt-in)."""

    def __init__(self, center: str = "grid"):
        super().__init__()
        self.center_mode = _lib.CENTER_MODES[center]

    def forward(self, volume, R):
        return VolumeRotationFunction.apply(volume, R, self.center_mode)


class VolumeConvolutionFunction(_NoBackward, torch.autograd.Function):
    @staticmethod
    def forward(ctx, v1, v2):
        v1 = _require_cuda_f32(v1, "input_volume1")
        v2 = _require_cuda_f32(v2, "input_volume2")
        if v1.shape != v2.shape or v1.dim() != 5 or not (v1.shape[2] == v1.shape[3] == v1.shape[4]):
            raise ValueError("expected two [B,C,L,L,L] volumes of equal shape")
        B, C, L = v1.shape[0], v1.shape[1], v1.shape[2]
        lib = _lib.get_lib()
        if not lib.dlpd_op_supported(L):
            raise _lib.DlpdError("VolumeConvolution: box size %d is not compiled into libdlpd" % L)
        out = torch.empty(B, C, 2 * L, 2 * L, 2 * L, device=v1.device, dtype=torch.float32)
        with torch.cuda.device(v1.device):
            _lib.check(lib, lib.dlpd_volume_correlate(_ptr(v1), _ptr(v2), _ptr(out), B, C, L, _stream_ptr(v1.device)))
        return out


class VolumeConvolution(nn.Module):
    """`TorchProteinLibrary.Volume.VolumeConvolution` (DockingModels.py:48,71; Docker.py:32,225):
    out[b,c,d mod 2L] = sum_r v1[b,c,r+d] v2[b,c,r].  `clip` is accepted and ignored in forward."""

    def __init__(self, clip=None):
        super().__init__()
        self.clip = clip

    def forward(self, input_volume1, input_volume2):
        return VolumeConvolutionFunction.apply(input_volume1, input_volume2)


class TypedCoords2Volume(nn.Module):
    """`TorchProteinLibrary.Volume.TypedCoords2Volume(box_size, resolution)` (Docker.py:31,204) in plain PyTorch.

    forward(coords [B,3*Nmax], num_atoms_of_type [B,T], offsets [B,T]) -> [B,T,L,L,L]; voxel i at i*resolution,
    every atom adds exp(-r^2/2) to the 5^3 voxels around floor(x/res) (parity unpinned)."""

    def __init__(self, box_size=120, resolution=1.0):
        super().__init__()
        self.box_size, self.resolution = int(box_size), float(resolution)

    def forward(self, input_coords_cpu, num_atoms_of_type, offsets):
        coords = input_coords_cpu
        B, T = num_atoms_of_type.shape
        L, res = self.box_size, self.resolution
        if coords.is_cuda and coords.dtype == torch.float32:
            return self._forward_cuda(coords, num_atoms_of_type, offsets)
        return self._forward_torch(coords, num_atoms_of_type, offsets)

    def _forward_cuda(self, coords, num_atoms_of_type, offsets):
        """One launch of the `ProjectTyped` kernel: one thread per atom slot, atomic scatter of the 5^3 stencil."""
        B, T = num_atoms_of_type.shape
        L = self.box_size
        c = coords.contiguous()
        nat = num_atoms_of_type.to(device=c.device, dtype=torch.int32).contiguous()
        off = offsets.to(device=c.device, dtype=torch.int32).contiguous()
        out = torch.empty(B, T, L, L, L, device=c.device, dtype=torch.float32)
        lib = _lib.get_lib()
        with torch.cuda.device(c.device):
            _lib.check(lib, lib.dlpd_project_typed(_ptr(c), _ptr(nat), _ptr(off), B, T, c.shape[1] // 3, L, self.resolution,
                                                  _ptr(out), _stream_ptr(c.device)))
        return out

    def _forward_torch(self, coords, num_atoms_of_type, offsets):
        B, T = num_atoms_of_type.shape
        L, res = self.box_size, self.resolution
        out = torch.zeros(B, T, L * L * L, device=coords.device, dtype=coords.dtype)
        nat = num_atoms_of_type.to("cpu")
        off = offsets.to("cpu")
        d = torch.arange(-2, 3, device=coords.device)
        stencil = torch.stack(torch.meshgrid(d, d, d, indexing="ij"), dim=-1).reshape(-1, 3)       # [125,3]
        for b in range(B):
            xyz = coords[b].reshape(-1, 3)
            types = torch.full((xyz.shape[0],), -1, dtype=torch.int64, device=coords.device)
            for t in range(T):
                o, n = int(off[b, t]), int(nat[b, t])
                types[o:o + n] = t
            keep = types >= 0
            xyz, types = xyz[keep], types[keep]
            base = torch.floor(xyz / res).to(torch.int64)
            ijk = base[:, None, :] + stencil[None, :, :]                                           # [Nat,125,3]
            ok = ((ijk >= 0) & (ijk < L)).all(dim=-1)
            r2 = ((xyz[:, None, :] - ijk.to(coords.dtype) * res) ** 2).sum(dim=-1)
            val = torch.exp(-r2 / 2.0) * ok.to(coords.dtype)
            c = ijk.clamp(0, L - 1)
            lin = (c[..., 0] * L + c[..., 1]) * L + c[..., 2]
            out[b].index_put_((types[:, None].expand_as(lin).reshape(-1), lin.reshape(-1)), val.reshape(-1), accumulate=True)
        return out.reshape(B, T, L, L, L)


def project_clash_rotated(lig_coords: torch.Tensor, R: torch.Tensor, box_size: int, resolution: float) -> torch.Tensor:
    """The in-loop ligand re-projection of `dockSE3` (Docker.py:221-224) for a batch of rotations, one CUDA launch:
    lig_coords [Nat,3] float64 (bbox-centred), R [n,3,3] float64, both on the GPU -> [n,L,L,L] float32 summed density
    of the rotated, box-centred atoms (`project(...).sum(dim=1)`)."""
    if not (lig_coords.is_cuda and R.is_cuda):
        raise _lib.DlpdError("project_clash_rotated needs CUDA tensors: libdlpd has no CPU path")
    c = lig_coords.to(torch.float64).contiguous()
    Rd = R.to(torch.float64).contiguous().reshape(-1, 9)
    n, L = Rd.shape[0], int(box_size)
    out = torch.empty(n, L, L, L, device=c.device, dtype=torch.float32)
    lib = _lib.get_lib()
    with torch.cuda.device(c.device):
        _lib.check(lib, lib.dlpd_project_clash(_ptr(c), _ptr(Rd), c.shape[0], n, L, float(resolution), _ptr(out),
                                              _stream_ptr(c.device)))
    return out
