"""GPU parity tests: the CUDA path (libdlpd.so through the package API) vs the CPU oracle and the golden fixtures.
Tolerances: indices exact; scores |s - s_ref| <= 1e-4 * max(|s_ref|, 1e-3 * max|V|) (BASELINE.json north_star)."""
import os

import pytest
import torch

from oracle import docking_oracle as O
from deeplocalproteindocking_b200 import (DockingSearch, VolumeConvolution, VolumeRotation, GlobalDockingModel,
                                          SimpleFilter, rotations, synthetic, _lib)

pytestmark = pytest.mark.gpu
R20 = rotations.Rotations(20).R
DEV = "cuda:0"


def _need_experimental():
    """The rejected epilogue variants and the mma.sync filter exist only in -DDLPD_EXPERIMENTAL=1 builds of libdlpd.so."""
    if not _lib.has_experimental():
        pytest.skip("libdlpd.so was built without -DDLPD_EXPERIMENTAL=1")


def _d(t):
    return t.double()


@pytest.mark.parametrize("L,center", [(8, "box"), (16, "grid"), (40, "box"), (80, "box")])
def test_volume_rotation_matches_oracle(L, center):
    g = torch.Generator().manual_seed(L)
    vol = torch.randn(2, 3, L, L, L, generator=g)
    R = R20[[17, 1234]]
    got = VolumeRotation(center)(vol.to(DEV), R.float().to(DEV)).cpu()
    want = O.rotate_volume(_d(vol), R, center)
    assert (got.double() - want).abs().max() < 5e-5
    # inputs untouched, new tensor returned, identity rotation is exact
    ident = VolumeRotation(center)(vol.to(DEV), torch.eye(3, device=DEV).repeat(2, 1, 1)).cpu()
    assert torch.equal(ident, vol)


@pytest.mark.parametrize("L", [4, 8, 10, 16, 20, 40, 80])
def test_volume_convolution_matches_oracle(L):
    g = torch.Generator().manual_seed(100 + L)
    B, C = (2, 3) if L <= 40 else (1, 2)
    v1 = torch.randn(B, C, L, L, L, generator=g)
    v2 = torch.randn(B, C, L, L, L, generator=g)
    got = VolumeConvolution(clip=5.0)(v1.to(DEV), v2.to(DEV)).cpu()
    want = O.correlate(_d(v1), _d(v2))
    assert got.shape == want.shape
    assert (got.double() - want).abs().max() <= 2e-5 * want.abs().max()
    assert got[:, :, L].abs().max() <= 2e-5 * want.abs().max()          # the d = -L plane has no overlap


def test_volume_convolution_golden_multiply_volumes(golden_dir):
    """Values produced by the reference's MultiplyVolumes.multiply (tests/golden/correlation.pt); the 6^3 fixture is
    zero-padded to a compiled size (same sums for |d| < 6)."""
    g = torch.load(os.path.join(golden_dir, "correlation.pt"), weights_only=False)
    pad = lambda v: torch.nn.functional.pad(v.float(), (0, 2, 0, 2, 0, 2))
    corr = VolumeConvolution()(pad(g["v1"])[None].to(DEV), pad(g["v2"])[None].to(DEV)).cpu()[0]
    for d, val in zip(g["d"].tolist(), g["values"]):
        t = [a % 16 for a in d]
        assert torch.allclose(corr[:, t[0], t[1], t[2]].double(), val, rtol=0, atol=2e-5 * float(g["values"].abs().max()))


def test_global_docking_model_golden(golden_dir):
    """Reference GlobalDockingModel.forward output through (a) the module-level drop-in, (b) the fused search."""
    g = torch.load(os.path.join(golden_dir, "global_model.pt"), weights_only=False)
    V = g["V_global"]
    filt = SimpleFilter([4, 6])
    filt.fc[0].weight.data, filt.fc[0].bias.data = g["W1"].float(), g["b1"].float()
    filt.fc[2].weight.data, filt.fc[2].bias.data = g["W2"].float(), g["b2"].float()
    model = GlobalDockingModel(torch.nn.Identity(), filt).to(DEV)
    with torch.no_grad():
        Vm = model([v.float().to(DEV) for v in g["rec"]], [v.float().to(DEV) for v in g["lig"]]).cpu()[0]
    assert (Vm.double() - V).abs().max() <= 1e-4 * V.abs().max()
    eng = DockingSearch(8, (4, 6), 5, max_rot=2, device=DEV)
    eng.set_receptor([v[0] for v in g["rec"]], torch.zeros(8, 8, 8))
    eng.set_ligand([v[0] for v in g["lig"]], torch.zeros(8, 8, 8))
    eng.set_filter(filt)
    best, scores = eng.search(torch.eye(3)[None], 1.0, return_scores=True)
    assert (scores[0].cpu().double() - V).abs().max() <= 1e-4 * V.abs().max()
    s, x, y, z = O.best_pose(V)
    assert [int(v) for v in best[0, 1:].cpu()] == [x, y, z]
    eng.close()


def _case(L, C0, C1, seed):
    g = torch.Generator().manual_seed(seed)
    L1 = L // 2
    rec = [torch.randn(C0, L, L, L, generator=g), torch.randn(C1, L1, L1, L1, generator=g)]
    lig = [torch.randn(C0, L, L, L, generator=g), torch.randn(C1, L1, L1, L1, generator=g)]
    rf = torch.rand(L, L, L, generator=g) * 2
    lf = torch.rand(L, L, L, generator=g) * 2
    return rec, lig, rf, lf


def _check_against_oracle(best, scores, rec, lig, rf, lf, Rs, filt, tau, clash_rot=None):
    spectra = O.ReceptorSpectra([_d(r)[None] for r in rec], _d(rf)[None, None])
    f64 = filt.to(torch.float64)
    for i in range(Rs.shape[0]):
        if clash_rot is None:
            V = O.score_rotation(spectra, [_d(l)[None] for l in lig], _d(lf), Rs[i], f64, tau)
            lrot = O.rotate_volume(_d(lf)[None, None], Rs[i][None])
        else:
            V = O.score_rotation(spectra, [_d(l)[None] for l in lig], _d(clash_rot[i]), Rs[i], f64, tau,
                                 ligand_forbidden_is_rotated=True)
            lrot = _d(clash_rot[i])[None, None]
        scale = float(V.abs().max())
        s, x, y, z = O.best_pose(V)
        assert [int(v) for v in best[i, 1:]] == [x, y, z], (i, best[i], (s, x, y, z))
        assert abs(float(best[i, 0]) - s) <= 1e-4 * max(abs(s), 1e-3 * scale)
        if scores is not None:
            clash = O.correlate(_d(rf)[None, None], lrot)[0, 0]
            safe = (clash - tau).abs() > 1e-4 * clash.abs().max()      # threshold-straddling voxels may flip in fp32
            assert (scores[i].double() - V).abs()[safe].max() <= 1e-4 * scale


@pytest.mark.parametrize("L,C0,C1,H,n", [(8, 4, 6, 5, 5), (16, 4, 6, 5, 4), (20, 4, 6, 5, 3), (8, 16, 32, 24, 3),
                                         (16, 16, 32, 24, 2), (20, 16, 32, 24, 5), (16, 32, 64, 48, 2), (20, 32, 64, 48, 3)])
def test_search_small_matches_oracle(L, C0, C1, H, n):
    rec, lig, rf, lf = _case(L, C0, C1, seed=L * 10 + C0)
    filt = O.FilterWeights.random(C0 + C1, H, seed=3)
    Rs = R20[torch.arange(n) * 311 + 5]
    tau = float(torch.quantile(O.correlate(rf[None, None], lf[None, None]).flatten(), 0.7))
    eng = DockingSearch(L, (C0, C1), H, max_rot=4, device=DEV)
    eng.set_receptor(rec, rf); eng.set_ligand(lig, lf); eng.set_filter((filt.W1, filt.b1, filt.W2, filt.b2))
    best, scores = eng.search(Rs, tau, return_scores=True)
    best_host = eng.search_host(Rs.float(), tau)
    eng.close()
    assert torch.equal(best.cpu(), best_host)
    _check_against_oracle(best.cpu(), scores.cpu(), rec, lig, rf, lf, Rs, filt, tau)


def test_search_full_size_80_matches_oracle():
    """BASELINE config shape: L=80, 16@80^3 + 32@40^3 + clash, 48->24->1 filter; 3 rotations (odd tail) vs fp64 oracle."""
    L, C0, C1 = 80, 16, 32
    rec, rec_clash = synthetic.learned_volumes(L, (C0, C1), seed=11)
    lig, lig_clash = synthetic.learned_volumes(L, (C0, C1), seed=12)
    filt = O.FilterWeights.random(C0 + C1, seed=13)
    Rs = R20[[7, 777, 1777]]
    tau = 30.0
    eng = DockingSearch(L, (C0, C1), max_rot=4, device=DEV)
    eng.set_receptor(rec, rec_clash); eng.set_ligand(lig, lig_clash); eng.set_filter((filt.W1, filt.b1, filt.W2, filt.b2))
    best, scores = eng.search(Rs, tau, return_scores=True)
    best2, _ = eng.search(Rs, tau)
    eng.close()
    assert torch.equal(best, best2)
    _check_against_oracle(best.cpu(), scores.cpu(), rec, lig, rec_clash, lig_clash, Rs, filt, tau)


@pytest.mark.parametrize("L", [80, 128])
def test_search_full_size_translation_equivariance(L):
    """Size-independent property at the BASELINE box sizes (no oracle needed): moving the ligand by an (even) voxel offset s
    inside its box moves every score by the same offset, V'[d] = V[d + s], so the best pose moves by -s and keeps its score.
    With R = identity the trilinear gather is exact, so the check isolates FFT correlation + filter + mask + argmin."""
    C0, C1, N = 16, 32, 2 * L
    rec, rec_clash = synthetic.learned_volumes(L, (C0, C1), seed=51)
    lig, lig_clash = synthetic.learned_volumes(L, (C0, C1), seed=52)
    q = L // 4                                       # ligand content confined to the central half of the box
    m0 = torch.zeros(L, L, L); m0[q:3 * q, q:3 * q, q:3 * q] = 1.0
    m1 = torch.zeros(L // 2, L // 2, L // 2); m1[q // 2:3 * q // 2, q // 2:3 * q // 2, q // 2:3 * q // 2] = 1.0
    lig = [lig[0] * m0, lig[1] * m1]
    lig_clash = lig_clash * m0
    s = (4, -6, 8)
    lig_s = [torch.roll(lig[0], s, dims=(1, 2, 3)), torch.roll(lig[1], tuple(v // 2 for v in s), dims=(1, 2, 3))]
    clash_s = torch.roll(lig_clash, s, dims=(0, 1, 2))
    filt = O.FilterWeights.random(C0 + C1, seed=53)
    eye = torch.eye(3)[None]
    eng = DockingSearch(L, (C0, C1), max_rot=2, device=DEV)
    eng.set_receptor(rec, rec_clash); eng.set_filter((filt.W1, filt.b1, filt.W2, filt.b2))
    out = {}
    for tau in (1e30, 30.0):                          # without and with the clash mask
        eng.set_ligand(lig, lig_clash)
        b0, v0 = eng.search(eye, tau, return_scores=True)
        eng.set_ligand(lig_s, clash_s)
        b1, v1 = eng.search(eye, tau, return_scores=True)
        out[tau] = (b0.cpu()[0], b1.cpu()[0], v0[0], v1[0])
    eng.close()
    idx = torch.arange(N, device=DEV)
    dd = torch.where(idx < L, idx, idx - N)
    ok1 = (dd.abs() < L - 10)                         # neither d nor d + s crosses the d = +-L seam of the circular index
    ok = ok1[:, None, None] & ok1[None, :, None] & ok1[None, None, :]
    for tau, (b0, b1, v0, v1) in out.items():
        want = torch.roll(v0, tuple(-a for a in s), dims=(0, 1, 2))      # want[d] = v0[d + s]
        scale = float(v0.abs().max())
        bad = ((v1 - want).abs() > 2e-5 * scale) & ok
        if tau > 1e29:
            assert int(bad.sum()) == 0
        else:
            # a voxel whose clash value sits on the threshold may flip with the rounding of the shifted transform
            assert int(bad.sum()) <= 1e-4 * float(ok.sum())
            assert float((v0 == 0).float().mean()) > 1e-4                # the mask is active
        assert scale > 0 and float((v1 - v0).abs().max()) > 1e-3 * scale  # the shift did something
        d0 = [int(b0[i + 1]) - (N if b0[i + 1] >= L else 0) for i in range(3)]
        if tau > 1e29 and all(abs(d0[i]) < L - 20 for i in range(3)):
            d1 = [int(b1[i + 1]) - (N if b1[i + 1] >= L else 0) for i in range(3)]
            assert d1 == [d0[i] - s[i] for i in range(3)]
            assert abs(float(b1[0] - b0[0])) <= 2e-5 * scale


def test_search_prerotated_clash_and_chunking_invariance():
    L, C0, C1, H = 20, 16, 32, 24
    rec, lig, rf, lf = _case(L, C0, C1, seed=5)
    filt = O.FilterWeights.random(C0 + C1, H, seed=6)
    Rs = R20[torch.arange(7) * 199 + 1]
    clash_rot = torch.stack([O.rotate_volume(lf[None, None], Rs[i][None].float())[0, 0] for i in range(7)])
    tau = 60.0
    out = []
    for max_rot in (2, 3, 16):
        eng = DockingSearch(L, (C0, C1), H, max_rot=max_rot, device=DEV)
        eng.set_receptor(rec, rf); eng.set_ligand(lig, None); eng.set_filter((filt.W1, filt.b1, filt.W2, filt.b2))
        with pytest.raises(_lib.DlpdError):
            eng.search(Rs, tau)                                     # no clash volume anywhere: loud failure
        b, _ = eng.search(Rs, tau, ligand_forbidden_rotated=clash_rot)
        out.append(b.cpu())
        eng.close()
    _check_against_oracle(out[0], None, rec, lig, rf, lf, Rs, filt, tau, clash_rot=clash_rot)
    for b in out[1:]:
        assert torch.equal(b[:, 1:], out[0][:, 1:])
        assert torch.allclose(b[:, 0], out[0][:, 0], rtol=1e-5, atol=1e-6)


def test_masked_ties_resolve_to_first_voxel():
    """Everything forbidden (tau below every clash value) -> V == 0 everywhere -> argmin is voxel (0,0,0) (Docker.py:91-96)."""
    L, C0, C1, H = 8, 4, 6, 5
    rec, lig, rf, lf = _case(L, C0, C1, seed=9)
    filt = O.FilterWeights.random(C0 + C1, H, seed=1)
    eng = DockingSearch(L, (C0, C1), H, device=DEV)
    eng.set_receptor(rec, rf); eng.set_ligand(lig, lf); eng.set_filter((filt.W1, filt.b1, filt.W2, filt.b2))
    best, scores = eng.search(R20[:3], -1.0, return_scores=True)
    eng.close()
    assert torch.count_nonzero(scores) == 0
    assert torch.equal(best.cpu(), torch.zeros(3, 4))


def test_errors_are_loud():
    with pytest.raises(_lib.DlpdError):
        DockingSearch(12, (4, 6), 5, device=DEV)                    # not a compiled configuration
    with pytest.raises(_lib.DlpdError):
        VolumeConvolution()(torch.zeros(1, 1, 8, 8, 8), torch.zeros(1, 1, 8, 8, 8))   # CPU tensors: no CPU path
    with pytest.raises(_lib.DlpdError):
        VolumeConvolution()(torch.zeros(1, 1, 12, 12, 12, device=DEV), torch.zeros(1, 1, 12, 12, 12, device=DEV))
    eng = DockingSearch(8, (4, 6), 5, device=DEV)
    with pytest.raises(_lib.DlpdError):
        eng.search(R20[:2], 1.0)                                    # nothing set yet
    eng.close()
    v = torch.zeros(1, 1, 8, 8, 8, device=DEV, requires_grad=True)
    out = VolumeConvolution()(v, v.detach())
    with pytest.raises(NotImplementedError):
        out.sum().backward()


# ---- the kernels of the 128^3 configuration: cluster plane (distributed shared memory; 8 CTAs at 128^3, 4 below) + 2x1 epilogue tile ----
@pytest.fixture
def cluster_mode():
    _lib.set_option("plane_mode", "cluster")
    _lib.set_option("epilogue_tile", 1)
    yield
    _lib.set_option("plane_mode", "auto")
    _lib.set_option("epilogue_tile", 0)


@pytest.mark.parametrize("L", [8, 16, 20, 40, 80])
def test_cluster_plane_convolution_matches_oracle(L, cluster_mode):
    test_volume_convolution_matches_oracle(L)


@pytest.mark.parametrize("L,C0,C1,H,n", [(16, 4, 6, 5, 4), (20, 16, 32, 24, 5)])
def test_cluster_plane_search_small_matches_oracle(L, C0, C1, H, n, cluster_mode):
    test_search_small_matches_oracle(L, C0, C1, H, n)


def test_cluster_plane_search_80_identical_poses(cluster_mode):
    """Same L=80 problem through both plane kernels: identical best poses, scores equal to rounding."""
    L, C0, C1 = 80, 16, 32
    rec, rec_clash = synthetic.learned_volumes(L, (C0, C1), seed=11)
    lig, lig_clash = synthetic.learned_volumes(L, (C0, C1), seed=12)
    filt = O.FilterWeights.random(C0 + C1, seed=13)
    Rs = R20[[7, 777, 1777, 42, 1000]]
    out = []
    for mode, tile in (("cluster", 1), ("cluster", 2), ("single", 2)):
        _lib.set_option("plane_mode", mode)
        _lib.set_option("epilogue_tile", tile)
        eng = DockingSearch(L, (C0, C1), max_rot=4, device=DEV)
        eng.set_receptor(rec, rec_clash); eng.set_ligand(lig, lig_clash); eng.set_filter((filt.W1, filt.b1, filt.W2, filt.b2))
        out.append(eng.search(Rs, 30.0)[0].cpu())
        eng.close()
    for b in out[:2]:
        assert torch.equal(b[:, 1:], out[2][:, 1:])
        assert torch.allclose(b[:, 0], out[2][:, 0], rtol=2e-5, atol=1e-6)


def test_volume_convolution_128_matches_oracle():
    L = 128
    g = torch.Generator().manual_seed(128)
    v1 = torch.randn(1, 1, L, L, L, generator=g)
    v2 = torch.randn(1, 1, L, L, L, generator=g)
    got = VolumeConvolution()(v1.to(DEV), v2.to(DEV)).cpu()
    want = O.correlate(_d(v1), _d(v2))
    assert (got.double() - want).abs().max() <= 2e-5 * want.abs().max()
    vol = torch.randn(1, 2, L, L, L, generator=g)
    R = R20[[321]]
    rot = VolumeRotation()(vol.to(DEV), R.float().to(DEV)).cpu()
    assert (rot.double() - O.rotate_volume(_d(vol), R)).abs().max() < 2e-4      # fp32 source coordinates: error grows with L


def test_search_full_size_128_matches_oracle():
    """BASELINE config 5 shape: L=128, 16@128^3 + 32@64^3 + clash; 2 rotations (one packed pair) vs the fp64 oracle."""
    L, C0, C1 = 128, 16, 32
    rec, rec_clash = synthetic.learned_volumes(L, (C0, C1), seed=21)
    lig, lig_clash = synthetic.learned_volumes(L, (C0, C1), seed=22)
    filt = O.FilterWeights.random(C0 + C1, seed=23)
    Rs = R20[[9, 1234]]
    tau = 30.0
    eng = DockingSearch(L, (C0, C1), max_rot=2, device=DEV)
    eng.set_receptor(rec, rec_clash); eng.set_ligand(lig, lig_clash); eng.set_filter((filt.W1, filt.b1, filt.W2, filt.b2))
    best, _ = eng.search(Rs, tau)
    eng.close()
    _check_against_oracle(best.cpu(), None, rec, lig, rec_clash, lig_clash, Rs, filt, tau)


def test_project_clash_matches_oracle():
    """CUDA re-projection of the rotated ligand atoms (Docker.py:221-224) vs the oracle, atoms partly outside the box."""
    from deeplocalproteindocking_b200.volume import project_clash_rotated
    g = torch.Generator().manual_seed(4)
    L, res = 20, 1.25
    coords = (torch.rand(400, 3, generator=g, dtype=torch.float64) - 0.5) * 27.0
    Rs = R20[[0, 77, 1500, 1853]]
    got = project_clash_rotated(coords.to(DEV), Rs.to(DEV), L, res).cpu()
    for i in range(4):
        want = O.ligand_clash_volume(coords, Rs[i], L, res)
        assert (got[i] - want).abs().max() <= 2e-6 * max(1.0, float(want.abs().max()))
    with pytest.raises(_lib.DlpdError):
        project_clash_rotated(coords, Rs, L, res)                      # CPU tensors: no CPU path


# ---- pipelined epilogue (load / transform / filter stages overlapped with named barriers) --------------------------
@pytest.fixture
def pipelined_epilogue():
    _need_experimental()
    _lib.set_option("epilogue_kind", "pipelined")
    yield
    _lib.set_option("epilogue_kind", "auto")


@pytest.mark.parametrize("L,C0,C1,H,n", [(8, 4, 6, 5, 5), (16, 4, 6, 5, 4), (20, 4, 6, 5, 3), (8, 16, 32, 24, 3),
                                         (16, 16, 32, 24, 2), (20, 16, 32, 24, 5)])
def test_pipelined_epilogue_small_matches_oracle(L, C0, C1, H, n, pipelined_epilogue):
    test_search_small_matches_oracle(L, C0, C1, H, n)


def test_pipelined_epilogue_80_matches_oracle(pipelined_epilogue):
    test_search_full_size_80_matches_oracle()


# ---- two half-tile epilogue CTAs per SM --------------------------------------------------------------------------
@pytest.fixture
def duo_epilogue():
    _need_experimental()
    _lib.set_option("epilogue_kind", "duo")
    yield
    _lib.set_option("epilogue_kind", "auto")


@pytest.mark.parametrize("L,C0,C1,H,n", [(8, 4, 6, 5, 5), (16, 4, 6, 5, 4), (20, 4, 6, 5, 3), (8, 16, 32, 24, 3),
                                         (16, 16, 32, 24, 2), (20, 16, 32, 24, 5)])
def test_duo_epilogue_small_matches_oracle(L, C0, C1, H, n, duo_epilogue):
    test_search_small_matches_oracle(L, C0, C1, H, n)


def test_duo_epilogue_80_matches_oracle(duo_epilogue):
    test_search_full_size_80_matches_oracle()


# ---- fused epilogue -------------------------------------------------------------------------------------------------
@pytest.fixture
def fused_epilogue():
    _need_experimental()
    _lib.set_option("epilogue_kind", "fused")
    yield
    _lib.set_option("epilogue_kind", "auto")


@pytest.mark.parametrize("L,C0,C1,H,n", [(8, 4, 6, 5, 5), (16, 4, 6, 5, 4), (20, 4, 6, 5, 3), (8, 16, 32, 24, 3),
                                         (16, 16, 32, 24, 2), (20, 16, 32, 24, 5)])
def test_fused_epilogue_small_matches_oracle(L, C0, C1, H, n, fused_epilogue):
    test_search_small_matches_oracle(L, C0, C1, H, n)


def test_fused_epilogue_80_matches_oracle(fused_epilogue):
    test_search_full_size_80_matches_oracle()


def test_epilogue_kinds_80_many_rotations_identical():
    """33 rotations (17 pairs, odd tail) at L=80 through all four epilogues: identical best poses and scores."""
    _need_experimental()
    L, C0, C1 = 80, 16, 32
    rec, rec_clash = synthetic.learned_volumes(L, (C0, C1), seed=31)
    lig, lig_clash = synthetic.learned_volumes(L, (C0, C1), seed=32)
    filt = O.FilterWeights.random(C0 + C1, seed=33)
    Rs = R20[torch.arange(33) * 53 + 2]
    out = []
    for kind in ("classic", "pipelined", "duo", "fused"):
        _lib.set_option("epilogue_kind", kind)
        _lib.set_option("filter_kind", "fp32")
        eng = DockingSearch(L, (C0, C1), max_rot=16, device=DEV)
        eng.set_receptor(rec, rec_clash); eng.set_ligand(lig, lig_clash); eng.set_filter((filt.W1, filt.b1, filt.W2, filt.b2))
        out.append(eng.search(Rs, 30.0)[0].cpu())
        eng.close()
    _lib.set_option("epilogue_kind", "auto")
    _lib.set_option("filter_kind", "auto")
    assert torch.equal(out[0], out[1]) and torch.equal(out[0], out[2]) and torch.equal(out[0], out[3])


# ---- first filter layer: packed FFMA chain vs fp16-split mma.sync -------------------------------------------------
@pytest.fixture
def tensor_filter():
    _need_experimental()
    _lib.set_option("filter_kind", "tensor")
    yield
    _lib.set_option("filter_kind", "auto")


@pytest.mark.parametrize("L,C0,C1,H,n", [(8, 16, 32, 24, 3), (16, 16, 32, 24, 2), (20, 16, 32, 24, 5)])
def test_tensor_filter_small_matches_oracle(L, C0, C1, H, n, tensor_filter):
    test_search_small_matches_oracle(L, C0, C1, H, n)


def test_tensor_filter_80_matches_oracle(tensor_filter):
    test_search_full_size_80_matches_oracle()


@pytest.mark.parametrize("scale", [1.0, 1e-3, 3e3])
def test_filter_kinds_80_agree(scale):
    """tcgen05 (3 x TF32 split, TMEM) vs FFMA2 first layer at L=80, 9 rotations, volumes scaled over six orders of magnitude
    (TF32 keeps fp32's exponent range: no scaling needed) and channels of very different size: identical best
    translations, scores and the whole masked score volume of one rotation within 1e-5 of the score range.  Experimental
    builds also run the mma.sync fp16-split variant."""
    L, C0, C1 = 80, 16, 32
    rec, rec_clash = synthetic.learned_volumes(L, (C0, C1), seed=41)
    lig, lig_clash = synthetic.learned_volumes(L, (C0, C1), seed=42)
    rec = [v * scale for v in rec]
    rec[0][3] *= 37.0                                   # channels of very different magnitude
    rec[0][7] *= 1.0 / 64.0
    filt = O.FilterWeights.random(C0 + C1, seed=43)
    W1 = filt.W1.clone()
    W1[:, :C0] /= scale                                 # keep the hidden pre-activations O(1)
    Rs = R20[torch.arange(9) * 201 + 5]
    out, vol = [], []
    kinds = ("fp32", "tcgen05") + (("tensor",) if _lib.has_experimental() else ())
    for kind in kinds:
        _lib.set_option("filter_kind", kind)
        eng = DockingSearch(L, (C0, C1), max_rot=8, device=DEV)
        eng.set_receptor(rec, rec_clash); eng.set_ligand(lig, lig_clash); eng.set_filter((W1, filt.b1, filt.W2, filt.b2))
        out.append(eng.search(Rs, 30.0)[0].cpu())
        vol.append(eng.search(Rs[:1], 30.0, return_scores=True)[1].cpu())
        eng.close()
    _lib.set_option("filter_kind", "auto")
    rng = float(vol[0].abs().max())
    for k in range(1, len(kinds)):
        assert torch.equal(out[0][:, 1:], out[k][:, 1:]), kinds[k]
        assert (out[0][:, 0] - out[k][:, 0]).abs().max() <= 1e-5 * rng, kinds[k]
        assert (vol[0] - vol[k]).abs().max() <= 1e-5 * rng, kinds[k]


@pytest.fixture(params=["fp32", "tcgen05"])
def filter_kind(request):
    _lib.set_option("filter_kind", request.param)
    yield request.param
    _lib.set_option("filter_kind", "auto")


@pytest.mark.parametrize("L,C0,C1,H,n", [(8, 16, 32, 24, 3), (16, 16, 32, 24, 2), (20, 16, 32, 24, 5), (20, 32, 64, 48, 3)])
def test_filter_kind_small_matches_oracle(L, C0, C1, H, n, filter_kind):
    """Both shipped forms of the first filter layer at sizes with partly filled warpgroups (64 / 128 / 160 translations per tile)."""
    test_search_small_matches_oracle(L, C0, C1, H, n)


def test_filter_kind_80_matches_oracle(filter_kind):
    test_search_full_size_80_matches_oracle()


def test_filter_kind_2x1_tile_matches_oracle(filter_kind):
    """The 2x1 epilogue tile (what 128^3 uses) at L = 20 and L = 80: 2 / 3 warpgroups of translations per tile."""
    _lib.set_option("epilogue_tile", 1)
    try:
        test_search_small_matches_oracle(20, 16, 32, 24, 5)
        test_search_full_size_80_matches_oracle()
    finally:
        _lib.set_option("epilogue_tile", 0)


# ---- two tcgen05 half-tile epilogue CTAs per SM ---------------------------------------------------------------------
@pytest.fixture
def tc_duo_epilogue():
    _need_experimental()
    _lib.set_option("epilogue_kind", "tc_duo")
    yield
    _lib.set_option("epilogue_kind", "auto")


@pytest.mark.parametrize("L,C0,C1,H,n", [(8, 16, 32, 24, 3), (16, 16, 32, 24, 2), (20, 16, 32, 24, 5)])
def test_tc_duo_epilogue_small_matches_oracle(L, C0, C1, H, n, tc_duo_epilogue):
    test_search_small_matches_oracle(L, C0, C1, H, n)


def test_tc_duo_epilogue_80_matches_oracle(tc_duo_epilogue):
    test_search_full_size_80_matches_oracle()


def test_tc_duo_epilogue_80_many_rotations_same_poses():
    """33 rotations (odd tail) at L = 80: the two-CTA epilogue and the default one give the same poses, scores to rounding."""
    _need_experimental()
    L, C0, C1 = 80, 16, 32
    rec, rec_clash = synthetic.learned_volumes(L, (C0, C1), seed=31)
    lig, lig_clash = synthetic.learned_volumes(L, (C0, C1), seed=32)
    filt = O.FilterWeights.random(C0 + C1, seed=33)
    Rs = R20[torch.arange(33) * 53 + 2]
    out = []
    for kind in ("auto", "tc_duo"):
        _lib.set_option("epilogue_kind", kind)
        eng = DockingSearch(L, (C0, C1), max_rot=16, device=DEV)
        eng.set_receptor(rec, rec_clash); eng.set_ligand(lig, lig_clash); eng.set_filter((filt.W1, filt.b1, filt.W2, filt.b2))
        out.append(eng.search(Rs, 30.0)[0].cpu())
        eng.close()
    _lib.set_option("epilogue_kind", "auto")
    assert torch.equal(out[0][:, 1:], out[1][:, 1:])
    assert torch.allclose(out[0][:, 0], out[1][:, 0], rtol=1e-5, atol=1e-6)


def test_search_80_wide_model_matches_oracle():
    """The representation classes' own default width (multiplier = 16: 32@80^3 + 64@40^3, H = 48,
    ProteinRepresentationModels.py:24,79): 2x1 epilogue tile, ONE level-1 column buffer refilled by a bulk copy per tile,
    160 TMEM columns per warpgroup.  Two rotations (one packed pair) vs the fp64 oracle."""
    L, C0, C1 = 80, 32, 64
    rec, rec_clash = synthetic.learned_volumes(L, (C0, C1), seed=71)
    lig, lig_clash = synthetic.learned_volumes(L, (C0, C1), seed=72)
    filt = O.FilterWeights.random(C0 + C1, seed=73)
    Rs = R20[[5, 1500]]
    eng = DockingSearch(L, (C0, C1), max_rot=2, device=DEV)
    eng.set_receptor(rec, rec_clash); eng.set_ligand(lig, lig_clash); eng.set_filter((filt.W1, filt.b1, filt.W2, filt.b2))
    best, _ = eng.search(Rs, 30.0)
    eng.close()
    _check_against_oracle(best.cpu(), None, rec, lig, rec_clash, lig_clash, Rs, filt, 30.0)


# ---- pipelined epilogue with the tcgen05 scoring stage (experimental builds) -----------------------------------------
def test_pipe_tc_epilogue_80_matches_oracle_and_default():
    _need_experimental()
    _lib.set_option("epilogue_kind", "pipe_tc")
    try:
        test_search_full_size_80_matches_oracle()
    finally:
        _lib.set_option("epilogue_kind", "auto")
    L, C0, C1 = 80, 16, 32
    rec, rec_clash = synthetic.learned_volumes(L, (C0, C1), seed=31)
    lig, lig_clash = synthetic.learned_volumes(L, (C0, C1), seed=32)
    filt = O.FilterWeights.random(C0 + C1, seed=33)
    Rs = R20[torch.arange(33) * 53 + 2]
    out = []
    for kind in ("auto", "pipe_tc"):
        _lib.set_option("epilogue_kind", kind)
        eng = DockingSearch(L, (C0, C1), max_rot=16, device=DEV)
        eng.set_receptor(rec, rec_clash); eng.set_ligand(lig, lig_clash); eng.set_filter((filt.W1, filt.b1, filt.W2, filt.b2))
        out.append(eng.search(Rs, 30.0)[0].cpu())
        eng.close()
    _lib.set_option("epilogue_kind", "auto")
    assert torch.equal(out[0][:, 1:], out[1][:, 1:])
    assert torch.allclose(out[0][:, 0], out[1][:, 0], rtol=1e-5, atol=1e-6)
