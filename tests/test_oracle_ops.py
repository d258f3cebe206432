"""Self-consistency of the oracle's unpinned operators (documented choices)."""
import math

import torch

from oracle import docking_oracle as O
from deeplocalproteindocking_b200 import rotations as rot


def _rand_rot(seed):
    a = rot.soi_euler_angles(20)[seed * 37 % 1854]
    return O.euler_to_matrix(*a)


def test_rotate_gather_equals_grid_sample():
    g = torch.Generator().manual_seed(3)
    for L, center in ((9, "box"), (12, "box"), (12, "grid")):
        vol = torch.randn(2, 3, L, L, L, generator=g, dtype=torch.float64)
        R = torch.stack([_rand_rot(1), _rand_rot(2)])
        a = O.rotate_volume(vol, R, center)
        b = O.rotate_volume_grid_sample(vol, R, center)
        assert torch.allclose(a, b, atol=1e-12)


def test_identity_rotation_is_identity():
    vol = torch.randn(1, 2, 10, 10, 10, dtype=torch.float64)
    out = O.rotate_volume(vol, torch.eye(3, dtype=torch.float64)[None])
    assert torch.equal(out, vol)


def test_rotation_direction_blob():
    """P4: a blob at c+u ends up at c+R.u (same direction as CoordsRotate r' = R r, Docker.py:218 vs :221)."""
    L = 32
    c = O.rotation_center(L)                       # centre of the default convention
    u = torch.tensor([6.0, -3.0, 4.0], dtype=torch.float64)
    ax = torch.arange(L, dtype=torch.float64)
    X, Y, Z = torch.meshgrid(ax, ax, ax, indexing="ij")

    def blob(p):
        return torch.exp(-((X - p[0]) ** 2 + (Y - p[1]) ** 2 + (Z - p[2]) ** 2) / 4.0)

    R = _rand_rot(5)
    out = O.rotate_volume(blob(c + u)[None, None], R[None])[0, 0]
    want = blob(c + R @ u)
    assert (out * want).sum() / math.sqrt((out * out).sum() * (want * want).sum()) > 0.995


def test_rotated_projection_consistent_with_volume_rotation():
    """center='box' keeps `rotate(project(atoms))` aligned with `project(rotate(atoms))` (unpinned (a))."""
    L, res = 32, 1.25
    g = torch.Generator().manual_seed(9)
    atoms = (torch.rand(300, 3, generator=g, dtype=torch.float64) - 0.5) * 14.0
    R = _rand_rot(8)
    v0 = O.ligand_clash_volume(atoms, torch.eye(3, dtype=torch.float64), L, res)
    v_rot = O.rotate_volume(v0[None, None].double(), R[None], center="box")[0, 0]
    v_proj = O.ligand_clash_volume(atoms, R, L, res).double()
    sim = (v_rot * v_proj).sum() / torch.sqrt((v_rot ** 2).sum() * (v_proj ** 2).sum())
    assert sim > 0.97
    v_grid = O.rotate_volume(v0[None, None].double(), R[None], center="grid")[0, 0]
    sim_grid = (v_grid * v_proj).sum() / torch.sqrt((v_grid ** 2).sum() * (v_proj ** 2).sum())
    assert sim > sim_grid


def test_default_rotation_is_affine_grid_plus_grid_sample():
    """The default centre ('grid') is what `F.affine_grid` + `F.grid_sample` do -- the construction TorchProteinLibrary's
    VolumeRotation is built on -- for BOTH align_corners settings (the centre is (L-1)/2 in index space either way):
    theta = P R^T P with P the (x,y,z) <-> (z,y,x) axis swap of grid_sample's coordinate order."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(3)
    for L in (9, 12):
        vol = torch.randn(2, 3, L, L, L, generator=g, dtype=torch.float64)
        R = torch.stack([_rand_rot(11), _rand_rot(12)])
        want = O.rotate_volume(vol, R)                                  # default centre
        P = torch.tensor([[0., 0., 1.], [0., 1., 0.], [1., 0., 0.]], dtype=torch.float64)
        theta = torch.zeros(2, 3, 4, dtype=torch.float64)
        theta[:, :, :3] = P @ R.transpose(1, 2) @ P
        for ac in (True, False):
            grid = F.affine_grid(theta, list(vol.shape), align_corners=ac)
            got = F.grid_sample(vol, grid, mode="bilinear", padding_mode="zeros", align_corners=ac)
            assert (got - want).abs().max() < 1e-9, (L, ac)
        assert (O.rotate_volume(vol, R, center="box") - want).abs().max() > 1e-3     # the opt-in convention differs


def test_dock_small_end_to_end_fp32_vs_fp64():
    g = torch.Generator().manual_seed(1)
    L = 8
    rec = [torch.rand(1, 4, L, L, L, generator=g), torch.rand(1, 6, L // 2, L // 2, L // 2, generator=g)]
    lig = [torch.rand(1, 4, L, L, L, generator=g), torch.rand(1, 6, L // 2, L // 2, L // 2, generator=g)]
    rf = torch.rand(1, 1, L, L, L, generator=g)
    lf = torch.rand(L, L, L, generator=g)
    filt = O.FilterWeights.random(10, seed=4)
    Rs = rot.Rotations(20).R[::400]
    b32, top32 = O.dock(rec, rf, lig, lf, Rs, filt, threshold=3.0, K=5)
    d = lambda t: t.double()
    b64, top64 = O.dock([d(r) for r in rec], d(rf), [d(l) for l in lig], d(lf), Rs, filt.to(torch.float64), threshold=3.0, K=5)
    assert torch.equal(b32[:, 1:], b64[:, 1:])
    assert torch.allclose(b32[:, 0], b64[:, 0], rtol=1e-4, atol=1e-6)
    assert [t[:4] for t in top32] == [t[:4] for t in top64]
