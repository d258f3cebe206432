"""Reader side of the on-disk pose format: the `.dat` files `Docker.write_conformations` produces
(reference writer /root/reference/src/Docker/Docker.py:107-133, 13 columns: the 3x3 rotation row-major, the translation in
Angstrom, the score) as the reference's consumer reads and applies them
(/root/reference/scripts/Results/Benchmark/DockerParser.py:35-69 `parse_output`, :77-87 `transform_ligand`).
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch

from .tpl import CoordsRotate, CoordsTranslate, getBBox

Pose = Tuple[torch.Tensor, torch.Tensor, float]           # (R [1,3,3] float64, t [1,3] float64, score)


def parse_output(filename: str, truncate_translation: bool = False) -> Optional[List[Pose]]:
    """Poses of one target in file order (best first).  Returns None when the file does not exist, like the reference.

    `truncate_translation=True` reproduces the reference reader literally: it stores `int(float(column))`
    (DockerParser.py:62-64) although the writer emits multiples of the 1.25 A resolution (Docker.py:122,132) -- a reference
    quirk; the default keeps the value that was written."""
    if not os.path.exists(filename):
        return None
    poses: List[Pose] = []
    with open(filename) as fin:
        for line in fin:
            s = line.split()
            if not s:
                continue
            if len(s) != 13:
                raise Exception("pose line with %d columns in %s (13 expected)" % (len(s), filename))
            v = [float(x) for x in s]
            R = torch.tensor(v[:9], dtype=torch.float64).reshape(1, 3, 3)
            t = torch.tensor([float(int(x)) if truncate_translation else x for x in v[9:12]], dtype=torch.float64).reshape(1, 3)
            poses.append((R, t, v[12]))
    return poses


def centre_on_bbox(coords: torch.Tensor, num_atoms: torch.Tensor) -> torch.Tensor:
    """DockerParser.load_protein (DockerParser.py:71-75): translate a structure so that its bounding box is centred at 0 --
    the frame in which the poses are expressed (the ligand is bbox-centred before rotation, Docker.py:54-59)."""
    a, b = getBBox(coords, num_atoms)
    return CoordsTranslate()(coords, -(a + b) * 0.5, num_atoms)


def transform_ligand(coords: torch.Tensor, num_atoms: torch.Tensor, pose: Pose) -> torch.Tensor:
    """DockerParser.transform_ligand (DockerParser.py:77-87): rotate the bbox-centred ligand by R, then translate by t."""
    R, t, _ = pose
    return CoordsTranslate()(CoordsRotate()(coords, R, num_atoms), t, num_atoms)
