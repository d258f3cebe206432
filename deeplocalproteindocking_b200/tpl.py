"""Import-compatible stand-in for the TorchProteinLibrary module surface the reference's search path imports
(/root/reference/src/Docker/Docker.py:8-10, src/Models/DockingModels.py:9-15, src/Utils/VolumeViz.py:10).

`install()` registers these names in `sys.modules`, so that the reference's own `Docker.py` / `DockingModels.py` run
unmodified on the sm_100a operators:

    TorchProteinLibrary.Volume         VolumeRotation, VolumeConvolution  -> libdlpd.so (dlpd_volume_rotate / _correlate)
                                       TypedCoords2Volume                 -> plain PyTorch (north_star: outside the loop)
    TorchProteinLibrary.FullAtomModel  CoordsTransform.{CoordsRotate, CoordsTranslate}, CoordsRotate, CoordsTranslate,
                                       getBBox, getRandomRotation          -> float64 CPU tensors like TorchProteinLibrary's
                                       PDB2CoordsUnordered, Coords2TypedCoords -> SYNTHETIC stand-ins: a "file name" is the
                                       name of a registered synthetic protein (`register_synthetic`) or a spec
                                       "synthetic:<n_atoms>:<radius>:<seed>"; they return tensors of the same shapes and
                                       dtypes as TorchProteinLibrary's (PDB parsing and the 11-type typing table are outside
                                       the search path: a real path raises with that message)
    _Volume                            Volume2Xplor (debug writer, VolumeViz.py:31) -> X-PLOR text map

Conventions follow the call sites: coords [B, 3*Nmax] float64 on the CPU, num_atoms [B] int, R [B,3,3], T [B,3];
`rotate(coords, R, num_atoms)` gives r' = R r (Docker.py:163,221), `translate` adds T to every real atom (Docker.py:150,222);
atoms beyond num_atoms[b] are padding and stay untouched.
"""
from __future__ import annotations

import sys
import types

import torch
import torch.nn as nn

from .volume import VolumeRotation, VolumeConvolution, TypedCoords2Volume


def _valid_mask(coords: torch.Tensor, num_atoms: torch.Tensor) -> torch.Tensor:
    B, n = coords.shape[0], coords.shape[1] // 3
    return torch.arange(n)[None, :] < num_atoms.reshape(B, 1).to(torch.int64)          # [B, Nmax]


class CoordsRotate(nn.Module):
    """TorchProteinLibrary.FullAtomModel.CoordsTransform.CoordsRotate: r' = R_b r for the first num_atoms[b] atoms."""

    def forward(self, coords: torch.Tensor, R: torch.Tensor, num_atoms: torch.Tensor) -> torch.Tensor:
        B = coords.shape[0]
        xyz = coords.reshape(B, -1, 3)
        rot = torch.einsum("bij,bnj->bni", R.to(coords.dtype).reshape(B, 3, 3), xyz)
        keep = _valid_mask(coords, num_atoms)[..., None]
        return torch.where(keep, rot, xyz).reshape(B, -1)


class CoordsTranslate(nn.Module):
    """TorchProteinLibrary.FullAtomModel.CoordsTranslate: r' = r + T_b for the first num_atoms[b] atoms."""

    def forward(self, coords: torch.Tensor, T: torch.Tensor, num_atoms: torch.Tensor) -> torch.Tensor:
        B = coords.shape[0]
        xyz = coords.reshape(B, -1, 3)
        keep = _valid_mask(coords, num_atoms)[..., None]
        return torch.where(keep, xyz + T.to(coords.dtype).reshape(B, 1, 3), xyz).reshape(B, -1)


def getBBox(coords: torch.Tensor, num_atoms: torch.Tensor):
    """(min corner [B,3], max corner [B,3]) over the real atoms of every structure (Docker.py:54)."""
    B = coords.shape[0]
    xyz = coords.reshape(B, -1, 3)
    keep = _valid_mask(coords, num_atoms)[..., None]
    big = torch.finfo(coords.dtype).max
    a = torch.where(keep, xyz, torch.full_like(xyz, big)).min(dim=1).values
    b = torch.where(keep, xyz, torch.full_like(xyz, -big)).max(dim=1).values
    return a, b


def getRandomRotation(batch_size: int) -> torch.Tensor:
    """[B,3,3] float64 rotations, uniform on SO(3) (unit quaternions from a normal sample), Docker.py:44."""
    q = torch.randn(batch_size, 4, dtype=torch.float64)
    q = q / q.norm(dim=1, keepdim=True)
    w, x, y, z = q.unbind(1)
    return torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], 1),
        torch.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], 1),
        torch.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1)], 1)


# ---- synthetic stand-ins for the PDB reader and the atom typer (Docker.py:37-38, 51-52) ----------------------------
NUM_ATOM_TYPES = 11
_SYNTHETIC = {}


def register_synthetic(name: str, coords: torch.Tensor, types: torch.Tensor) -> None:
    """Make `name` loadable by `PDB2CoordsUnordered`: coords [Nat,3] Angstrom, types [Nat] in 0..10."""
    _SYNTHETIC[name] = (torch.as_tensor(coords, dtype=torch.float64).reshape(-1, 3).clone(),
                        torch.as_tensor(types, dtype=torch.int64).reshape(-1).clone())


def _synthetic_protein(name: str):
    if name in _SYNTHETIC:
        return _SYNTHETIC[name]
    if isinstance(name, str) and name.startswith("synthetic:"):
        from .synthetic import random_protein
        try:
            _, n, radius, seed = name.split(":")
            atoms = random_protein(int(n), float(radius), int(seed))
        except ValueError:
            raise Exception("bad synthetic protein spec (synthetic:<n_atoms>:<radius>:<seed>)", name)
        return atoms.coords, atoms.types
    raise NotImplementedError(
        "PDB2CoordsUnordered: %r is not a registered synthetic protein -- PDB parsing is outside the docking search path and "
        "is not provided by deeplocalproteindocking_b200 (use TorchProteinLibrary's reader, tpl.register_synthetic, or pass "
        "atoms directly to deeplocalproteindocking_b200.Docker)" % (name,))


class PDB2CoordsUnordered(nn.Module):
    """Stand-in with TorchProteinLibrary's return signature (Docker.py:51):
    `(coords [B,3 Nmax] float64, chains [B,Nmax] uint8, resnames [B,Nmax,4] uint8, resnums [B,Nmax] int32,
    atomnames [B,Nmax,4] uint8, num_atoms [B] int32)`, all on the CPU, padded with zeros.  The synthetic atom TYPE travels
    in `atomnames[..., 0]` (there are no real residue / atom names to type from)."""

    def forward(self, filenames):
        prots = [_synthetic_protein(f) for f in filenames]
        B, nmax = len(prots), max(p[0].shape[0] for p in prots)
        coords = torch.zeros(B, 3 * nmax, dtype=torch.float64)
        chains = torch.zeros(B, nmax, dtype=torch.uint8)
        resnames = torch.zeros(B, nmax, 4, dtype=torch.uint8)
        resnums = torch.zeros(B, nmax, dtype=torch.int32)
        atomnames = torch.zeros(B, nmax, 4, dtype=torch.uint8)
        num_atoms = torch.zeros(B, dtype=torch.int32)
        for b, (c, t) in enumerate(prots):
            n = c.shape[0]
            coords[b, :3 * n] = c.reshape(-1)
            chains[b, :n] = ord("A")
            resnums[b, :n] = torch.arange(n, dtype=torch.int32) // 8 + 1
            atomnames[b, :n, 0] = t.to(torch.uint8)
            num_atoms[b] = n
        return coords, chains, resnames, resnums, atomnames, num_atoms


class Coords2TypedCoords(nn.Module):
    """Stand-in with TorchProteinLibrary's return signature (Docker.py:52): atoms stably sorted by type,
    `(coords [B,3 Nmax] float64, num_atoms_of_type [B,11] int32, offsets [B,11] int32)`."""

    def forward(self, coords, resnames, atomnames, num_atoms):
        B, nmax = coords.shape[0], coords.shape[1] // 3
        out = torch.zeros_like(coords)
        nat = torch.zeros(B, NUM_ATOM_TYPES, dtype=torch.int32)
        off = torch.zeros(B, NUM_ATOM_TYPES, dtype=torch.int32)
        for b in range(B):
            n = int(num_atoms[b])
            t = atomnames[b, :n, 0].to(torch.int64)
            if n and int(t.max()) >= NUM_ATOM_TYPES:
                raise Exception("atom type out of range", int(t.max()))
            order = torch.argsort(t, stable=True)
            out[b, :3 * n] = coords[b].reshape(nmax, 3)[:n][order].reshape(-1)
            counts = torch.bincount(t, minlength=NUM_ATOM_TYPES)[:NUM_ATOM_TYPES]
            nat[b] = counts.to(torch.int32)
            off[b] = (torch.cumsum(counts, 0) - counts).to(torch.int32)
        return out, nat, off


def Volume2Xplor(volume: torch.Tensor, filename: str, resolution: float) -> None:
    """`_Volume.Volume2Xplor(volume_cpu [L,L,L], filename, resolution)` (VolumeViz.py:31): X-PLOR / CNS text density map,
    voxel i at i * resolution Angstrom."""
    v = volume.detach().to("cpu", torch.float32)
    if v.dim() != 3:
        raise ValueError("Volume2Xplor expects a [L,L,L] volume")
    nx, ny, nz = v.shape
    with open(filename, "w") as f:
        f.write("\n       1 !NTITLE\n REMARKS deeplocalproteindocking_b200 density map\n")
        f.write("%8d%8d%8d%8d%8d%8d%8d%8d%8d\n" % (nx, 0, nx - 1, ny, 0, ny - 1, nz, 0, nz - 1))
        f.write("%12.5E%12.5E%12.5E%12.5E%12.5E%12.5E\n" % (nx * resolution, ny * resolution, nz * resolution, 90.0, 90.0, 90.0))
        f.write("ZYX\n")
        for k in range(nz):
            f.write("%8d\n" % k)
            plane = v[:, :, k].t().reshape(-1).tolist()                  # x fastest, then y
            for i in range(0, len(plane), 6):
                f.write("".join("%12.5E" % x for x in plane[i:i + 6]) + "\n")
        f.write("%8d\n" % -9999)
        f.write("%12.4E %12.4E\n" % (float(v.mean()), float(v.std())))


def install(force: bool = True) -> None:
    """Register the stand-in modules under the names the reference imports.  With force=False an already importable
    TorchProteinLibrary is left alone."""
    if not force and "TorchProteinLibrary" in sys.modules:
        return

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    tpl = mod("TorchProteinLibrary")
    tpl.Volume = mod("TorchProteinLibrary.Volume", VolumeRotation=VolumeRotation, VolumeConvolution=VolumeConvolution,
                     TypedCoords2Volume=TypedCoords2Volume)
    ct = mod("TorchProteinLibrary.FullAtomModel.CoordsTransform", CoordsRotate=CoordsRotate, CoordsTranslate=CoordsTranslate,
             getBBox=getBBox, getRandomRotation=getRandomRotation)
    tpl.FullAtomModel = mod("TorchProteinLibrary.FullAtomModel", CoordsTransform=ct, CoordsRotate=CoordsRotate,
                            CoordsTranslate=CoordsTranslate, PDB2CoordsUnordered=PDB2CoordsUnordered,
                            Coords2TypedCoords=Coords2TypedCoords, getBBox=getBBox, getRandomRotation=getRandomRotation)
    mod("_Volume", Volume2Xplor=Volume2Xplor)
