"""CPU ORACLE for the docking search -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this file.  Nothing under `deeplocalproteindocking_b200/` does.

It restates, in plain CPU PyTorch (`torch.fft`), the arithmetic of the reference's
exhaustive search `Docker.dockSE3` (`/root/reference/src/Docker/Docker.py:184-238`).

Parity status (see DESIGN.md "Oracle"):
  * PINNED against the reference's own code, imported unmodified in the authoring container
    (`oracle/make_golden.py` -> `tests/golden/*.pt`):  `Rotations.writeMatrix`,
    `MultiplyVolumes.multiply` (defines the correlation), `GlobalDockingModel.forward`,
    `LocalDockingModel.forward`, `SimpleFilter`, `Docker.update_top`, `Docker.write_conformations`.
  * PARITY UNPINNED (source lives in TorchProteinLibrary @16166ce, not vendored, no tests):
    `VolumeRotation`'s rotation centre / boundary rule, `VolumeConvolution(clip=...)`,
    `TypedCoords2Volume`'s Gaussian.  The documented choices are in each docstring.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# rotations  (Rotations.py:14-32)
# --------------------------------------------------------------------------------------
def euler_to_matrix(phi: float, theta: float, psi: float) -> torch.Tensor:
    """One ZXZ rotation matrix, float64; follows `Rotations.writeMatrix` (Rotations.py:14-32)."""
    cpsi, spsi = math.cos(psi), math.sin(psi)
    ct, st = math.cos(theta), math.sin(theta)
    cphi, sphi = math.cos(phi), math.sin(phi)
    return torch.tensor(
        [[cpsi * cphi - spsi * ct * sphi, -cpsi * sphi - spsi * ct * cphi, spsi * st],
         [spsi * cphi + cpsi * ct * sphi, -spsi * sphi + cpsi * ct * cphi, -cpsi * st],
         [st * sphi, st * cphi, ct]], dtype=torch.float64)


# --------------------------------------------------------------------------------------
# VolumeRotation  (Docker.py:40,218)   -- PARITY UNPINNED: centre & boundary are our choice
# --------------------------------------------------------------------------------------
def rotation_center(L: int, center: str = "grid") -> float:
    """'grid' (DEFAULT): c = (L-1)/2 -- what `F.affine_grid` + `F.grid_sample` do for either `align_corners` setting
    (tests/test_oracle_ops.py::test_default_rotation_is_affine_grid_plus_grid_sample), i.e. the construction
    TorchProteinLibrary's VolumeRotation is built on;
    'box': c = L/2 (voxel i sits at i*res and atoms are centred at box_length/2, Docker.py:34-35) -- opt-in."""
    if center == "box":
        return L / 2.0
    if center == "grid":
        return (L - 1) / 2.0
    raise ValueError(center)


def rotate_volume(volume: torch.Tensor, R: torch.Tensor, center: str = "grid") -> torch.Tensor:
    """out[b,c,p] = trilinear(volume[b,c], ctr + R_b^T (p - ctr)), zero outside [0,L-1]^3.

    p = (x,y,z) = tensor dims (2,3,4).  volume [B,C,L,L,L]; R [B,3,3].  Computed in volume.dtype.
    Explicit 8-tap gather (no grid_sample) so the arithmetic is visible; `rotate_volume_grid_sample`
    is the equivalent call the survey names and the tests check the two against each other."""
    B, C, L = volume.shape[0], volume.shape[1], volume.shape[2]
    dt = volume.dtype
    ctr = rotation_center(L, center)
    ax = torch.arange(L, dtype=dt) - ctr
    px, py, pz = torch.meshgrid(ax, ax, ax, indexing="ij")
    P = torch.stack([px, py, pz], dim=-1).reshape(-1, 3)            # [L^3,3] (p - ctr)
    Rm = R.to(dt)
    # s = ctr + R^T (p-ctr)  ->  s_a = ctr + sum_b R[b,a] (p-ctr)_b  = ((p-ctr) @ R)[a]
    S = torch.einsum("nb,kba->kna", P, Rm) + ctr                    # [B,L^3,3]
    S0 = torch.floor(S)
    Fr = S - S0
    I0 = S0.to(torch.int64)
    out = torch.zeros(B, C, L * L * L, dtype=dt)
    flat = volume.reshape(B, C, L * L * L)
    for dx in (0, 1):
        wx = Fr[..., 0] if dx else 1 - Fr[..., 0]
        ix = I0[..., 0] + dx
        for dy in (0, 1):
            wy = Fr[..., 1] if dy else 1 - Fr[..., 1]
            iy = I0[..., 1] + dy
            for dz in (0, 1):
                wz = Fr[..., 2] if dz else 1 - Fr[..., 2]
                iz = I0[..., 2] + dz
                ok = (ix >= 0) & (ix < L) & (iy >= 0) & (iy < L) & (iz >= 0) & (iz < L)
                idx = (ix.clamp(0, L - 1) * L + iy.clamp(0, L - 1)) * L + iz.clamp(0, L - 1)
                w = (wx * wy * wz) * ok.to(dt)                       # [B,L^3]
                out += torch.gather(flat, 2, idx.unsqueeze(1).expand(B, C, -1)) * w.unsqueeze(1)
    return out.reshape(B, C, L, L, L)


def rotate_volume_grid_sample(volume: torch.Tensor, R: torch.Tensor, center: str = "grid") -> torch.Tensor:
    """Same operator through `F.grid_sample(mode='bilinear', padding_mode='zeros', align_corners=True)`."""
    B, C, L = volume.shape[0], volume.shape[1], volume.shape[2]
    dt = volume.dtype
    ctr = rotation_center(L, center)
    ax = torch.arange(L, dtype=dt) - ctr
    px, py, pz = torch.meshgrid(ax, ax, ax, indexing="ij")
    P = torch.stack([px, py, pz], dim=-1)                           # [L,L,L,3]
    S = torch.einsum("xyzb,kba->kxyza", P, R.to(dt)) + ctr          # source voxel coords (x,y,z)
    g = 2.0 * S / (L - 1) - 1.0
    grid = torch.stack([g[..., 2], g[..., 1], g[..., 0]], dim=-1)   # grid[...,0] addresses the LAST dim
    return F.grid_sample(volume, grid, mode="bilinear", padding_mode="zeros", align_corners=True)


# --------------------------------------------------------------------------------------
# VolumeConvolution  (DockingModels.py:48,71 ; Docker.py:32,225)
# --------------------------------------------------------------------------------------
def correlate(v1: torch.Tensor, v2: torch.Tensor) -> torch.Tensor:
    """[B,C,L,L,L] x2 -> [B,C,2L,2L,2L]:  out[c, d mod 2L] = sum_r v1[c,r+d] * v2[c,r].

    This is the quantity the reference's training path computes directly
    (`MultiplyVolumes.multiply`, MultiplyVolumes.py:13-47) and feeds to the same filter
    (DockingModels.py:115-119 vs :71,82), which pins sign, layout and normalisation.
    `clip` of `VolumeConvolution(clip=5.0)` is ignored in forward (unpinned)."""
    L = v1.shape[-1]
    N = 2 * L
    f1 = torch.fft.rfftn(v1, s=(N, N, N), dim=(-3, -2, -1))
    f2 = torch.fft.rfftn(v2, s=(N, N, N), dim=(-3, -2, -1))
    return torch.fft.irfftn(f1 * torch.conj(f2), s=(N, N, N), dim=(-3, -2, -1))


def correlate_direct(v1: torch.Tensor, v2: torch.Tensor, d: Sequence[int]) -> torch.Tensor:
    """sum_r v1[c,r+d] v2[c,r] for ONE integer translation d, by slicing; [C,L,L,L] -> [C].
    Own restatement of MultiplyVolumes.multiply (MultiplyVolumes.py:13-47) without the 8-case split."""
    L = v1.shape[-1]
    sl1, sl2 = [], []
    for a in d:
        a = int(a)
        if a >= 0:
            sl1.append(slice(a, L)); sl2.append(slice(0, L - a))
        else:
            sl1.append(slice(0, L + a)); sl2.append(slice(-a, L))
    return (v1[(slice(None), *sl1)] * v2[(slice(None), *sl2)]).sum(dim=(1, 2, 3))


# --------------------------------------------------------------------------------------
# GlobalDockingModel.forward  (DockingModels.py:63-84)  + SimpleFilter (DockingModels.py:23-37)
# --------------------------------------------------------------------------------------
class FilterWeights:
    """The SimpleFilter MLP: s = W2 . relu(W1 . v + b1) + b2  (DockingModels.py:28-32)."""

    def __init__(self, W1, b1, W2, b2):
        self.W1, self.b1, self.W2, self.b2 = W1, b1, W2.reshape(-1), b2.reshape(())

    @staticmethod
    def random(n_in: int, n_hidden: Optional[int] = None, seed: int = 0, dtype=torch.float32) -> "FilterWeights":
        """xavier_uniform weights like `init_weights` (DockingModels.py:17-21), default nn.Linear biases."""
        n_hidden = n_in // 2 if n_hidden is None else n_hidden
        g = torch.Generator().manual_seed(seed)
        a1 = math.sqrt(6.0 / (n_in + n_hidden))
        a2 = math.sqrt(6.0 / (n_hidden + 1))
        W1 = (torch.rand(n_hidden, n_in, generator=g, dtype=torch.float64) * 2 - 1) * a1
        W2 = (torch.rand(1, n_hidden, generator=g, dtype=torch.float64) * 2 - 1) * a2
        b1 = (torch.rand(n_hidden, generator=g, dtype=torch.float64) * 2 - 1) / math.sqrt(n_in)
        b2 = (torch.rand(1, generator=g, dtype=torch.float64) * 2 - 1) / math.sqrt(n_hidden)
        return FilterWeights(W1.to(dtype), b1.to(dtype), W2.to(dtype), b2.to(dtype))

    def to(self, dtype) -> "FilterWeights":
        return FilterWeights(self.W1.to(dtype), self.b1.to(dtype), self.W2.to(dtype), self.b2.to(dtype))

    def __call__(self, v: torch.Tensor) -> torch.Tensor:
        """v [...,n_in] -> [...]"""
        h = torch.relu(v @ self.W1.t() + self.b1)
        return h @ self.W2 + self.b2


def upsample_nearest_circular(corr: torch.Tensor, size: int) -> torch.Tensor:
    """U[c,t] = corr[c, t // f] per axis (F.interpolate nearest, DockingModels.py:74-76)."""
    f = size // corr.shape[-1]
    if f == 1:
        return corr
    return corr.repeat_interleave(f, -3).repeat_interleave(f, -2).repeat_interleave(f, -1)


def global_score(receptor_volumes: List[torch.Tensor], ligand_volumes: List[torch.Tensor],
                 filt: FilterWeights) -> torch.Tensor:
    """[B,2L,2L,2L] score volume of `GlobalDockingModel.forward` (DockingModels.py:63-84):
    per-level correlation, nearest upsample to 2L, channel concat (level 0 first), per-voxel MLP."""
    N = 2 * receptor_volumes[0].shape[2]
    convs = [upsample_nearest_circular(correlate(r, l), N) for r, l in zip(receptor_volumes, ligand_volumes)]
    V = torch.cat(convs, dim=1).permute(0, 2, 3, 4, 1)
    return filt(V)


def masked_score(score: torch.Tensor, clash_corr: torch.Tensor, threshold: float) -> torch.Tensor:
    """V = (clash < tau).float() * score   (Docker.py:226,232).  Forbidden voxels become (+-)0.0."""
    return (clash_corr < threshold).to(score.dtype) * score


# --------------------------------------------------------------------------------------
# Docker.update_top  (Docker.py:86-105)
# --------------------------------------------------------------------------------------
def best_pose(V: torch.Tensor) -> Tuple[float, int, int, int]:
    """First extraction of update_top: hierarchical torch.min over z, y, x (Docker.py:91-96).
    Ties resolve to the first index along each axis = first voxel in row-major (x,y,z) order."""
    mz, iz = torch.min(V, dim=2)
    my, iy = torch.min(mz, dim=1)
    mx, ix = torch.min(my, dim=0)
    x = int(ix); y = int(iy[x]); z = int(iz[x, y])
    return float(V[x, y, z]), x, y, z


def update_top_literal(V: torch.Tensor, K: int) -> List[Tuple[int, int, int, float]]:
    """The K-iteration loop of Docker.update_top (Docker.py:88-98), verbatim semantics:
    take the minimum (first in row-major order), record it, overwrite it with 0.0, repeat.
    Mutates a copy.  O(K * volume): use on small volumes only."""
    V = V.clone()
    top = []
    for _ in range(K):
        s, x, y, z = best_pose(V)
        top.append((x, y, z, s))
        V[x, y, z] = 0.0
    return top


def update_top_fast(V: torch.Tensor, K: int) -> List[Tuple[int, int, int, float]]:
    """Same list as `update_top_literal` in O(volume log volume).

    While negative scores remain they come out in (score, row-major index) order.  Once the
    minimum is >= 0 the loop degenerates: extracted voxels are 0.0, so (a) if any voxel is <= 0 the
    first voxel (row-major) whose value is 0 -- among original zeros and already-extracted voxels --
    is returned again and again with score (+-)0.0; (b) if everything is positive the smallest
    positive is taken once, becomes 0.0, and is then returned for all remaining iterations."""
    N1, N2, N3 = V.shape
    flat = V.reshape(-1)
    top: List[Tuple[int, int, int, float]] = []

    def unravel(i):
        i = int(i)
        return i // (N2 * N3), (i // N3) % N2, i % N3

    neg = torch.nonzero(flat < 0).reshape(-1)
    if neg.numel():
        vals = flat[neg]
        order = torch.argsort(vals, stable=True)                    # ties keep row-major order
        for j in order[:K].tolist():
            top.append((*unravel(neg[j]), float(vals[j])))
    if len(top) == K:
        return top
    zero_candidates = torch.nonzero(flat == 0).reshape(-1).tolist()
    extracted = [(x * N2 + y) * N3 + z for x, y, z, _ in top]
    pool = zero_candidates + extracted
    if not pool:
        i = int(torch.argmin(flat))                                  # all positive: first minimum
        top.append((*unravel(i), float(flat[i])))
        pool = [i]
        if len(top) == K:
            return top
    i0 = min(pool)
    s0 = float(flat[i0]) if i0 in zero_candidates else 0.0          # keeps -0.0 of an original voxel
    first = True
    while len(top) < K:
        top.append((*unravel(i0), s0 if first else 0.0))
        first = False
    return top


def merge_top(top_list: list, new_items: list, K: int) -> list:
    """`self.top_list` update: append, stable sort by score, truncate (Docker.py:100-105)."""
    top_list = top_list + new_items
    top_list.sort(key=lambda t: t[4])
    return top_list[:K]


def index_to_translation(idx: Sequence[int], box_size: int, resolution: float) -> Tuple[float, float, float]:
    """Circular index -> signed translation in Angstrom (Docker.py:112-122)."""
    out = []
    for a in idx:
        t = float(a)
        if a >= box_size:
            t = -(2 * box_size - a)
        out.append(t * resolution)
    return tuple(out)


def format_conformation(R: torch.Tensor, t: Sequence[float], score: float) -> str:
    """One line of the pose file: 9 matrix entries, 3 translations, score; tab separated %f (Docker.py:129-133)."""
    vals = [float(R[i, j]) for i in range(3) for j in range(3)] + [float(x) for x in t]
    return "".join("%f\t" % v for v in vals) + "%f\n" % score


# --------------------------------------------------------------------------------------
# TypedCoords2Volume / coordinate ops  -- PARITY UNPINNED (TPL recollection)
# --------------------------------------------------------------------------------------
def project_atoms(coords: torch.Tensor, types: torch.Tensor, n_types: int, box_size: int,
                  resolution: float) -> torch.Tensor:
    """[Nat,3] Angstrom coords + [Nat] type ids -> [n_types,L,L,L] density (Docker.py:31,204).

    Voxel i sits at i*resolution; every atom adds exp(-|x - i*res|^2 / 2) to the 5^3 voxels within
    +-2 cells of floor(x/res) that fall inside the box."""
    L = box_size
    dt = coords.dtype
    vol = torch.zeros(n_types, L * L * L, dtype=dt)
    base = torch.floor(coords / resolution).to(torch.int64)          # [Nat,3]
    for dx in range(-2, 3):
        for dy in range(-2, 3):
            for dz in range(-2, 3):
                ijk = base + torch.tensor([dx, dy, dz])
                ok = ((ijk >= 0) & (ijk < L)).all(dim=1)
                r2 = ((coords - ijk.to(dt) * resolution) ** 2).sum(dim=1)
                val = torch.exp(-r2 / 2.0) * ok.to(dt)
                lin = (ijk[:, 0].clamp(0, L - 1) * L + ijk[:, 1].clamp(0, L - 1)) * L + ijk[:, 2].clamp(0, L - 1)
                vol.index_put_((types, lin), val, accumulate=True)
    return vol.reshape(n_types, L, L, L)


def ligand_clash_volume(lcoords: torch.Tensor, R: torch.Tensor, box_size: int, resolution: float) -> torch.Tensor:
    """Per-rotation ligand clash volume of dockSE3 (Docker.py:221-224): rotate the bbox-centred atoms
    (r' = R r, CPU float64), translate to the box centre box_length/2, project in float32, sum types."""
    c = box_size * resolution / 2.0
    x = (lcoords.to(torch.float64) @ R.to(torch.float64).t() + c).to(torch.float32)
    return project_atoms(x, torch.zeros(x.shape[0], dtype=torch.int64), 1, box_size, resolution)[0]


# --------------------------------------------------------------------------------------
# the search loop  (Docker.py:184-238)
# --------------------------------------------------------------------------------------
class ReceptorSpectra:
    """rfftn of the zero-padded receptor volumes, computed once (the reference recomputes per call)."""

    def __init__(self, receptor_volumes: List[torch.Tensor], receptor_forbidden: torch.Tensor):
        self.levels = []
        for v in receptor_volumes:
            N = 2 * v.shape[-1]
            self.levels.append((torch.fft.rfftn(v, s=(N, N, N), dim=(-3, -2, -1)), N))
        N = 2 * receptor_forbidden.shape[-1]
        self.forbidden = torch.fft.rfftn(receptor_forbidden, s=(N, N, N), dim=(-3, -2, -1))


def score_rotation(spectra: ReceptorSpectra, ligand_volumes: List[torch.Tensor], ligand_forbidden: torch.Tensor,
                   R: torch.Tensor, filt: FilterWeights, threshold: float, center: str = "grid",
                   ligand_forbidden_is_rotated: bool = False) -> torch.Tensor:
    """Masked score volume V [2L,2L,2L] of ONE rotation (Docker.py:218-232) with precomputed receptor spectra.

    ligand_volumes: list of [1,C_k,L_k,L_k,L_k]; ligand_forbidden [L,L,L] -- either the un-rotated
    summed density (rotated here with the same operator as the feature volumes) or, when
    `ligand_forbidden_is_rotated`, the re-projected rotated atoms (Docker.py:221-224)."""
    dt = ligand_volumes[0].dtype
    Rb = R.reshape(1, 3, 3).to(dt)
    N = 2 * ligand_volumes[0].shape[-1]
    convs = []
    for (spec, n), lig in zip(spectra.levels, ligand_volumes):
        rot = rotate_volume(lig, Rb, center)
        fl = torch.fft.rfftn(rot, s=(n, n, n), dim=(-3, -2, -1))
        convs.append(upsample_nearest_circular(torch.fft.irfftn(spec * torch.conj(fl), s=(n, n, n), dim=(-3, -2, -1)), N))
    score = filt(torch.cat(convs, dim=1).permute(0, 2, 3, 4, 1))[0]
    lf = ligand_forbidden.reshape(1, 1, *ligand_forbidden.shape[-3:])
    if not ligand_forbidden_is_rotated:
        lf = rotate_volume(lf, Rb, center)
    clash = torch.fft.irfftn(spectra.forbidden * torch.conj(torch.fft.rfftn(lf, s=(N, N, N), dim=(-3, -2, -1))),
                             s=(N, N, N), dim=(-3, -2, -1))[0, 0]
    return masked_score(score, clash, threshold)


def dock(receptor_volumes, receptor_forbidden, ligand_volumes, ligand_forbidden, Rs, filt, threshold,
         K: int = 1, center: str = "grid", ligand_forbidden_rotated=None):
    """Restatement of the dockSE3 loop over rotations `Rs [n,3,3]`.

    Returns (best [n,4] float64 rows (score,x,y,z), top_list of (rot,x,y,z,score) like Docker.top_list).
    `ligand_forbidden_rotated(i)` may supply the per-rotation clash volume (re-projected atoms)."""
    spectra = ReceptorSpectra(receptor_volumes, receptor_forbidden)
    best = torch.zeros(Rs.shape[0], 4, dtype=torch.float64)
    top_list: list = []
    for i in range(Rs.shape[0]):
        if ligand_forbidden_rotated is not None:
            V = score_rotation(spectra, ligand_volumes, ligand_forbidden_rotated(i), Rs[i], filt, threshold,
                               center, ligand_forbidden_is_rotated=True)
        else:
            V = score_rotation(spectra, ligand_volumes, ligand_forbidden, Rs[i], filt, threshold, center)
        s, x, y, z = best_pose(V)
        best[i] = torch.tensor([s, x, y, z], dtype=torch.float64)
        if K > 0:
            items = [(i, x, y, z, sc) for x, y, z, sc in update_top_fast(V, K)]
            top_list = merge_top(top_list, items, K)
    return best, top_list
