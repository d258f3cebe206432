"""Rotation grids for the exhaustive search.

Host-side mirror of the reference's `Rotations` class
(`/root/reference/src/Utils/Rotations.py:8-66`): same attribute (`.R`, float64 CPU
`[N,3,3]`), same Euler convention (`writeMatrix`, `Rotations.py:14-32`:
`R = Rz(psi) . Rx(theta) . Rz(phi)`), same sets (`data/oim*.eul`).

Differences, all documented in DESIGN.md:
  * the reference builds the file name with `'oim%d.eul'` so only 20/15/12/10 load
    (`Rotations.py:39`); here 8 and 6 load too and any explicit `.eul` path works;
  * the six published sets ship in factorised form (`deeplocalproteindocking_b200/data/soi_axes.json`, see
    `tools/pack_soi.py`) and are rebuilt to the exact decimal values of the tables;
  * `angle_inc=4`: `oim04.eul` is a missing blob in the reference snapshot, so -- like
    the reference, which raises "Can't find rotation angles" when the file is absent --
    `Rotations(4)` raises.  `Rotations(4, allow_surrogate=True)` (or
    `angle_inc="fibonacci4"`) opts into a deterministic SURROGATE of the same size:
    2578 Fibonacci-lattice axes x 90 in-plane angles = 232 020 rotations, same
    3-column format.  It is NOT the published 4-degree SOI set and is labelled as a
    surrogate wherever it is used (bench.py configs 3 and 5).
"""
from __future__ import annotations

import json
import math
import os
from functools import lru_cache

import numpy as np
import torch

_DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")      # package data: travels with an installed package
SOI_SIZES = {20: 1854, 15: 4392, 12: 8580, 10: 14868, 8: 29025, 6: 68760, 4: 232020}
_SURROGATE_AXES = {4: 2578}


@lru_cache(maxsize=None)
def _soi_axes():
    with open(os.path.join(_DATA_DIR, "soi_axes.json")) as f:
        return json.load(f)


def _fibonacci_axes(m: int) -> np.ndarray:
    """m quasi-uniform unit vectors -> (theta, psi) with R.z = (sin psi sin theta, -cos psi sin theta, cos theta)."""
    i = np.arange(m, dtype=np.float64) + 0.5
    nz = 1.0 - 2.0 * i / m
    az = i * (math.pi * (3.0 - math.sqrt(5.0)))
    s = np.sqrt(np.maximum(0.0, 1.0 - nz * nz))
    nx, ny = s * np.cos(az), s * np.sin(az)
    theta = np.arccos(np.clip(nz, -1.0, 1.0))
    psi = np.arctan2(nx, -ny)
    return np.stack([theta, psi], axis=1)


def soi_euler_angles(angle_inc: int, allow_surrogate: bool = False) -> np.ndarray:
    """`[N,3]` float64 (phi, theta, psi) radians, in the file order of `oim<inc>.eul`.

    Values are the 6-decimal numbers of the text tables (phi fastest, from -pi in
    steps of `angle_inc` degrees).  Sets that are not shipped raise unless `allow_surrogate`."""
    angle_inc = int(angle_inc)
    if 360 % angle_inc:
        raise Exception("Can't find rotation angles:", angle_inc)
    nphi = 360 // angle_inc
    phi = np.array([float("%f" % (-math.pi + k * angle_inc * math.pi / 180.0)) for k in range(nphi)])
    axes = _soi_axes().get(str(angle_inc))
    if axes is not None:
        tp = np.array([[float(t), float(p)] for t, p in axes], dtype=np.float64)
    elif angle_inc in _SURROGATE_AXES and allow_surrogate:
        tp = _fibonacci_axes(_SURROGATE_AXES[angle_inc])
        tp = np.array([[float("%f" % t), float("%f" % p)] for t, p in tp])
    else:
        raise Exception("Can't find rotation angles:", angle_inc)
    out = np.empty((tp.shape[0], nphi, 3), dtype=np.float64)
    out[:, :, 0] = phi[None, :]
    out[:, :, 1] = tp[:, None, 0]
    out[:, :, 2] = tp[:, None, 1]
    return out.reshape(-1, 3)


def read_eul(path: str) -> np.ndarray:
    """Parse a `phi theta psi` text table (whitespace separated, one rotation per line)."""
    rows = []
    with open(path, "r") as fin:
        for line in fin:
            s = line.split()
            if len(s) >= 3:
                rows.append((float(s[0]), float(s[1]), float(s[2])))
    if not rows:
        raise Exception("Can't find rotation angles:", path)
    return np.asarray(rows, dtype=np.float64)


def write_eul(path: str, angles: np.ndarray) -> None:
    """Emit the reference's text format (`'%f\\t %f\\t %f'` per line)."""
    with open(path, "w") as f:
        for phi, theta, psi in np.asarray(angles, dtype=np.float64):
            f.write("%f\t %f\t %f\n" % (phi, theta, psi))


def euler_to_matrix(angles) -> torch.Tensor:
    """ZXZ Euler angles `[N,3]` (phi, theta, psi) -> `[N,3,3]` float64.

    Element-for-element the formula of `Rotations.writeMatrix` (`Rotations.py:22-32`)."""
    a = torch.as_tensor(np.asarray(angles), dtype=torch.float64).reshape(-1, 3)
    phi, theta, psi = a[:, 0], a[:, 1], a[:, 2]
    cpsi, spsi = torch.cos(psi), torch.sin(psi)
    cth, sth = torch.cos(theta), torch.sin(theta)
    cphi, sphi = torch.cos(phi), torch.sin(phi)
    R = torch.empty(a.shape[0], 3, 3, dtype=torch.float64)
    R[:, 0, 0] = cpsi * cphi - spsi * cth * sphi
    R[:, 0, 1] = -cpsi * sphi - spsi * cth * cphi
    R[:, 0, 2] = spsi * sth
    R[:, 1, 0] = spsi * cphi + cpsi * cth * sphi
    R[:, 1, 1] = -spsi * sphi + cpsi * cth * cphi
    R[:, 1, 2] = -cpsi * sth
    R[:, 2, 0] = sth * sphi
    R[:, 2, 1] = sth * cphi
    R[:, 2, 2] = cth
    return R


class Rotations(object):
    """Drop-in for `src.Utils.Rotations` (`Rotations.py:8-66`): `.R` is `[N,3,3]` float64 on CPU.

    `angle_inc` may be one of 20/15/12/10/8/6 (shipped sets) or a path to an `.eul` file; 4 raises like the
    reference does for a missing table unless `allow_surrogate=True` / `angle_inc="fibonacci4"` asks for the labelled
    surrogate.  `verbose` keeps the reference's two print lines."""

    def __init__(self, angle_inc=12, verbose=False, allow_surrogate=False):
        self.surrogate = False
        if angle_inc == "fibonacci4":
            angle_inc, allow_surrogate = 4, True
        self.allow_surrogate = bool(allow_surrogate)
        self.loadSOI(angle_inc)
        if verbose:
            print("Angle increment:", angle_inc)
            print("Number of rotations:", self.R.size(0))

    def loadSOI(self, angle_inc):
        if isinstance(angle_inc, str):
            if not os.path.exists(angle_inc):
                raise Exception("Can't find rotation angles:", angle_inc)
            self.angles = read_eul(angle_inc)
        else:
            inc = int(angle_inc)
            if inc != angle_inc or inc not in SOI_SIZES:
                raise Exception("Can't find rotation angles:", angle_inc)
            self.angles = soi_euler_angles(inc, self.allow_surrogate)
            self.surrogate = inc in _SURROGATE_AXES
            assert self.angles.shape[0] == SOI_SIZES[inc]
        self.R = euler_to_matrix(self.angles)

    def __len__(self):
        return int(self.R.size(0))
