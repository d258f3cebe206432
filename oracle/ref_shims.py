"""`sys.modules` shims that let the reference's UNMODIFIED Python import in the authoring container.

TEST INFRASTRUCTURE.  Used by `oracle/make_golden.py` (and by tests that are skipped when
`/root/reference` is absent).  The reference imports TorchProteinLibrary, `_Volume` and se3cnn, none of
which is installed or vendored; `src/__init__.py` creates directories and asserts on the authors'
disks (`/root/reference/src/__init__.py:52-71`).  We inject:

  * a fake `src` package exposing REPOSITORY_DIR (pointing at a directory that holds `data/oim*.eul`)
    whose `__path__` is the real `/root/reference/src`, so `src.Utils`, `src.Models`, `src.Docker` are the
    reference's own files;
  * `TorchProteinLibrary.Volume` / `.FullAtomModel` modules whose operators are the ORACLE's
    (`docking_oracle.rotate_volume`, `.correlate`, `.project_atoms`) -- i.e. the unpinned TPL
    semantics are our documented choices, everything else executed is reference code;
  * empty stand-ins for `_Volume`, `se3cnn`, `matplotlib` names the modules import but the path never calls.
"""
from __future__ import annotations

import os
import sys
import types

import torch

from . import docking_oracle as O

REFERENCE_ROOT = "/root/reference"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "Docker"))


class _VolumeRotation(torch.nn.Module):
    def forward(self, volume, R):
        return O.rotate_volume(volume, R)


class _VolumeConvolution(torch.nn.Module):
    def __init__(self, clip=None):
        super().__init__()
        self.clip = clip

    def forward(self, v1, v2):
        return O.correlate(v1, v2)


class _TypedCoords2Volume(torch.nn.Module):
    def __init__(self, box_size=120, resolution=1.0):
        super().__init__()
        self.box_size, self.resolution = box_size, resolution

    def forward(self, coords, num_atoms_of_type, offsets):
        B, T = num_atoms_of_type.shape
        out = []
        for b in range(B):
            xyz = coords[b].reshape(-1, 3)
            types = torch.zeros(xyz.shape[0], dtype=torch.int64)
            keep = torch.zeros(xyz.shape[0], dtype=torch.bool)
            for t in range(T):
                o, n = int(offsets[b, t]), int(num_atoms_of_type[b, t])
                types[o:o + n] = t
                keep[o:o + n] = True
            out.append(O.project_atoms(xyz[keep], types[keep], T, self.box_size, self.resolution))
        return torch.stack(out)


class _CoordsRotate(torch.nn.Module):
    def forward(self, coords, R, num_atoms):
        B = coords.shape[0]
        xyz = coords.reshape(B, -1, 3)
        return torch.einsum("bij,bnj->bni", R.to(coords.dtype), xyz).reshape(B, -1)


class _CoordsTranslate(torch.nn.Module):
    def forward(self, coords, T, num_atoms):
        B = coords.shape[0]
        return (coords.reshape(B, -1, 3) + T.to(coords.dtype).reshape(B, 1, 3)).reshape(B, -1)


def _getBBox(coords, num_atoms):
    B = coords.shape[0]
    xyz = coords.reshape(B, -1, 3)
    return xyz.min(dim=1).values, xyz.max(dim=1).values


def _getRandomRotation(batch_size):
    g = torch.Generator().manual_seed(7)
    q = torch.randn(batch_size, 4, generator=g, dtype=torch.float64)
    q = q / q.norm(dim=1, keepdim=True)
    w, x, y, z = q.unbind(1)
    return torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], 1),
        torch.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], 1),
        torch.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1)], 1)


class _Unavailable(torch.nn.Module):
    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, *a, **k):
        raise RuntimeError("not available in the oracle shims")


def install(repository_dir: str):
    """Inject the shims; `repository_dir/data/oim*.eul` must exist.  Returns the reference modules
    (Rotations class, Models package, Docker class)."""
    if not reference_available():
        raise RuntimeError("reference not present")

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    tpl = mod("TorchProteinLibrary")
    tpl.Volume = mod("TorchProteinLibrary.Volume", VolumeRotation=_VolumeRotation, VolumeConvolution=_VolumeConvolution,
                     TypedCoords2Volume=_TypedCoords2Volume)
    ct = mod("TorchProteinLibrary.FullAtomModel.CoordsTransform", CoordsRotate=_CoordsRotate,
             CoordsTranslate=_CoordsTranslate)
    tpl.FullAtomModel = mod("TorchProteinLibrary.FullAtomModel", CoordsTransform=ct, CoordsRotate=_CoordsRotate,
                            CoordsTranslate=_CoordsTranslate, PDB2CoordsUnordered=_Unavailable,
                            Coords2TypedCoords=_Unavailable, getBBox=_getBBox, getRandomRotation=_getRandomRotation)
    mod("_Volume")
    se3 = mod("se3cnn", SE3Convolution=_Unavailable)
    se3.non_linearities = mod("se3cnn.non_linearities", ScalarActivation=_Unavailable)
    se3.blocks = mod("se3cnn.blocks", GatedBlock=_Unavailable)
    se3.filter = mod("se3cnn.filter", low_pass_filter=None)

    src = mod("src", REPOSITORY_DIR=repository_dir)
    src.__path__ = [os.path.join(REFERENCE_ROOT, "src")]
    import importlib
    Utils = importlib.import_module("src.Utils")
    Models = importlib.import_module("src.Models")
    DockerMod = importlib.import_module("src.Docker.Docker")
    return Utils.Rotations, Models, DockerMod.Docker
