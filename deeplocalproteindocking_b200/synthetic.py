"""Synthetic inputs for the BASELINE.json configurations: no datasets or checkpoints are available,
so proteins are random atom clouds and the representation network is randomly initialised."""
from __future__ import annotations

import math
from typing import List, Tuple

import torch

from .docker import Atoms, NUM_ATOM_TYPES


def random_protein(n_atoms: int, radius: float, seed: int) -> Atoms:
    """Points uniform in a ball of `radius` Angstrom, atom types uniform in 0..10."""
    g = torch.Generator().manual_seed(seed)
    v = torch.randn(n_atoms, 3, generator=g, dtype=torch.float64)
    v = v / v.norm(dim=1, keepdim=True)
    r = radius * torch.rand(n_atoms, 1, generator=g, dtype=torch.float64) ** (1.0 / 3.0)
    types = torch.randint(0, NUM_ATOM_TYPES, (n_atoms,), generator=g)
    return Atoms(v * r, types)


def config_pair(box_size: int = 80, seed: int = 1234) -> Tuple[Atoms, Atoms]:
    """Receptor 3000 atoms in a 28 A ball, ligand 1500 atoms in a 20 A ball at L=80 (scaled with the box)."""
    s = box_size / 80.0
    return (random_protein(int(3000 * s ** 3) or 8, 28.0 * s, seed),
            random_protein(int(1500 * s ** 3) or 8, 20.0 * s, seed + 1))


def learned_volumes(box_size: int, channels=(16, 32), seed: int = 0, device="cpu") -> Tuple[List[torch.Tensor], torch.Tensor]:
    """"Multi-channel learned volumes" generated directly (configs 4 and 5): zero-mean noise inside a ball,
    smoothed by one 3^3 box blur; returns ([C0,L^3], [C1,(L/2)^3]) and a positive clash density [L^3]."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for lvl, c in enumerate(channels):
        L = box_size >> lvl
        ax = torch.arange(L, dtype=torch.float32) - L / 2.0
        X, Y, Z = torch.meshgrid(ax, ax, ax, indexing="ij")
        ball = ((X * X + Y * Y + Z * Z) < (0.3 * L) ** 2).float()
        v = torch.randn(c, L, L, L, generator=g) * ball
        v = torch.nn.functional.avg_pool3d(v[None], 3, stride=1, padding=1)[0]
        out.append(v.to(device))
    L = box_size
    ax = torch.arange(L, dtype=torch.float32) - L / 2.0
    X, Y, Z = torch.meshgrid(ax, ax, ax, indexing="ij")
    clash = torch.rand(L, L, L, generator=g) * ((X * X + Y * Y + Z * Z) < (0.3 * L) ** 2).float() * 3.0
    return out, clash.to(device)
