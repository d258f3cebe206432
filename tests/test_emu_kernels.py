"""Kernel bodies (host emulation, see tests/emu/driver.py) vs the oracle.  CPU only."""
import os

import pytest
import torch

from oracle import docking_oracle as O
from deeplocalproteindocking_b200 import rotations as rot
from tests.emu import driver as emu

R20 = rot.Rotations(20).R


def _case(L, C0, C1, seed):
    g = torch.Generator().manual_seed(seed)
    L1 = L // 2
    # zero-mean features keep the filter well conditioned (all-positive volumes drive every hidden unit far
    # from 0 and the score is then a small difference of large numbers -- even the fp32 oracle misses 1e-4 there)
    rec = [torch.randn(1, C0, L, L, L, generator=g), torch.randn(1, C1, L1, L1, L1, generator=g)]
    lig = [torch.randn(1, C0, L, L, L, generator=g), torch.randn(1, C1, L1, L1, L1, generator=g)]
    rf = torch.rand(1, 1, L, L, L, generator=g) * 2
    lf = torch.rand(L, L, L, generator=g) * 2
    return rec, lig, rf, lf


@pytest.mark.parametrize("L", [4, 8, 10, 16, 20])
def test_emu_correlate_matches_oracle(L):
    g = torch.Generator().manual_seed(L)
    v1 = torch.randn(2, 2, L, L, L, generator=g)
    v2 = torch.randn(2, 2, L, L, L, generator=g)
    got = emu.volume_correlate(v1, v2)
    want = O.correlate(v1.double(), v2.double())
    assert (got.double() - want).abs().max() <= 2e-5 * want.abs().max()


@pytest.mark.parametrize("L,center", [(8, "box"), (10, "box"), (9, "grid"), (16, "grid")])
def test_emu_rotate_matches_oracle(L, center):
    g = torch.Generator().manual_seed(L)
    vol = torch.randn(2, 3, L, L, L, generator=g)
    R = R20[[100, 900]]
    got = emu.volume_rotate(vol, R, center)
    want = O.rotate_volume(vol.double(), R, center)
    assert (got.double() - want).abs().max() < 2e-5


@pytest.mark.parametrize("L,C0,C1,H,n", [(8, 4, 6, 5, 3), (16, 4, 6, 5, 2), (20, 4, 6, 5, 3), (8, 16, 32, 24, 2),
                                         (20, 16, 32, 24, 3)])
def test_emu_search_matches_oracle(L, C0, C1, H, n):
    rec, lig, rf, lf = _case(L, C0, C1, seed=L + C0)
    filt = O.FilterWeights.random(C0 + C1, H, seed=3)
    Rs = R20[torch.arange(n) * 311 + 5]
    tau = float(torch.quantile(O.correlate(rf, lf[None, None]).flatten(), 0.7))
    plan = emu.EmuPlan(L, C0, C1, H, max_rot=2)
    try:
        plan.set_receptor(rec[0][0], rec[1][0], rf[0, 0])
        plan.set_ligand(lig[0][0], lig[1][0], lf)
        plan.set_filter(filt.W1, filt.b1, filt.W2, filt.b2)
        best, scores = plan.search(Rs, tau, want_scores=True)
        best_h = plan.search_host(Rs, tau)
    finally:
        plan.close()
    assert torch.equal(best, best_h)
    d = lambda t: t.double()
    spectra = O.ReceptorSpectra([d(r) for r in rec], d(rf))
    f64 = filt.to(torch.float64)
    for i in range(n):
        V = O.score_rotation(spectra, [d(l) for l in lig], d(lf), Rs[i], f64, tau)
        scale = V.abs().max()
        # voxels whose clash value sits on the threshold may flip in fp32: compare away from it
        clash = O.correlate(d(rf), O.rotate_volume(d(lf)[None, None], Rs[i][None]))[0, 0]
        safe = (clash - tau).abs() > 1e-4 * clash.abs().max()
        assert ((scores[i].double() - V).abs()[safe]).max() <= 1e-4 * scale
        s, x, y, z = O.best_pose(V)
        assert [int(v) for v in best[i, 1:]] == [x, y, z]
        assert abs(float(best[i, 0]) - s) <= 1e-4 * max(abs(s), 1e-3 * float(scale))


def test_emu_search_with_prerotated_clash_and_odd_count():
    L, C0, C1, H = 8, 4, 6, 5
    rec, lig, rf, lf = _case(L, C0, C1, seed=77)
    filt = O.FilterWeights.random(C0 + C1, H, seed=5)
    Rs = R20[[3, 700, 1500]]
    clash_rot = torch.stack([O.rotate_volume(lf[None, None], Rs[i][None].float())[0, 0] for i in range(3)])
    tau = 4.0
    plan = emu.EmuPlan(L, C0, C1, H, max_rot=4)
    try:
        plan.set_receptor(rec[0][0], rec[1][0], rf[0, 0])
        plan.set_ligand(lig[0][0], lig[1][0], None)
        plan.set_filter(filt.W1, filt.b1, filt.W2, filt.b2)
        best_ext, _ = plan.search(Rs, tau, clash_rot=clash_rot)
        with pytest.raises(Exception):
            plan.search(Rs, tau)                       # no clash volume available -> loud failure
        plan.set_ligand(lig[0][0], lig[1][0], lf)
        best_int, _ = plan.search(Rs, tau)
    finally:
        plan.close()
    assert torch.equal(best_ext[:, 1:], best_int[:, 1:])
    assert torch.allclose(best_ext[:, 0], best_int[:, 0], rtol=1e-5, atol=1e-6)


def test_emu_global_model_golden(golden_dir):
    """Reference GlobalDockingModel.forward output (tests/golden/global_model.pt) through the kernel bodies."""
    g = torch.load(os.path.join(golden_dir, "global_model.pt"), weights_only=False)
    L, C0, C1, H = 8, 4, 6, 5
    plan = emu.EmuPlan(L, C0, C1, H, max_rot=2)
    try:
        plan.set_receptor(g["rec"][0][0], g["rec"][1][0], torch.zeros(L, L, L))
        plan.set_ligand(g["lig"][0][0], g["lig"][1][0], torch.zeros(L, L, L))
        plan.set_filter(g["W1"], g["b1"], g["W2"], g["b2"])
        best, scores = plan.search(torch.eye(3)[None], tau=1.0, want_scores=True)
    finally:
        plan.close()
    V = g["V_global"]
    assert (scores[0].double() - V).abs().max() <= 1e-4 * V.abs().max()
    s, x, y, z = O.best_pose(V)
    assert [int(v) for v in best[0, 1:]] == [x, y, z]


def test_emu_unsupported_configuration_fails_loudly():
    with pytest.raises(Exception):
        emu.EmuPlan(12, 4, 6, 5)
    with pytest.raises(Exception):
        emu.volume_correlate(torch.zeros(1, 1, 12, 12, 12), torch.zeros(1, 1, 12, 12, 12))


@pytest.fixture
def cluster_mode():
    """Force the cluster plane kernel (PlaneX, 4 CTAs at these sizes) and the 2x1 epilogue tile -- the kernels the 128^3 configuration uses."""
    emu.set_option(emu.OPT_PLANE_MODE, emu.PLANE_CLUSTER)
    emu.set_option(emu.OPT_EPILOGUE_TILE, 1)
    yield
    emu.set_option(emu.OPT_PLANE_MODE, emu.PLANE_AUTO)
    emu.set_option(emu.OPT_EPILOGUE_TILE, 0)


@pytest.mark.parametrize("L", [8, 16, 20])
def test_emu_cluster_plane_correlate_matches_oracle(L, cluster_mode):
    g = torch.Generator().manual_seed(L)
    v1 = torch.randn(1, 3, L, L, L, generator=g)
    v2 = torch.randn(1, 3, L, L, L, generator=g)
    got = emu.volume_correlate(v1, v2)
    want = O.correlate(v1.double(), v2.double())
    assert (got.double() - want).abs().max() <= 2e-5 * want.abs().max()


@pytest.mark.parametrize("L,C0,C1,H,n", [(8, 4, 6, 5, 3), (16, 4, 6, 5, 2), (20, 16, 32, 24, 3)])
def test_emu_cluster_plane_search_matches_oracle(L, C0, C1, H, n, cluster_mode):
    test_emu_search_matches_oracle(L, C0, C1, H, n)


@pytest.fixture
def cluster8_mode():
    """The 8-CTA cluster form of PlaneX (what N = 256 runs: S1 folded straight from global memory, radix-8 exchange), compiled
    for L = 16 so that the emulation can afford it (-DDLPD_PLANEX_CL_16=8)."""
    prev = emu.use_variant(["-DDLPD_PLANEX_CL_16=8"], "cl8")
    emu.set_option(emu.OPT_PLANE_MODE, emu.PLANE_CLUSTER)
    emu.set_option(emu.OPT_EPILOGUE_TILE, 1)
    yield
    emu.set_option(emu.OPT_PLANE_MODE, emu.PLANE_AUTO)
    emu.set_option(emu.OPT_EPILOGUE_TILE, 0)
    emu.restore(prev)


def test_emu_cluster8_plane_correlate_and_search_match_oracle(cluster8_mode):
    g = torch.Generator().manual_seed(16)
    v1 = torch.randn(1, 2, 16, 16, 16, generator=g)
    v2 = torch.randn(1, 2, 16, 16, 16, generator=g)
    want = O.correlate(v1.double(), v2.double())
    assert (emu.volume_correlate(v1, v2).double() - want).abs().max() <= 2e-5 * want.abs().max()
    test_emu_search_matches_oracle(16, 4, 6, 5, 2)
    test_emu_search_matches_oracle(16, 16, 32, 24, 2)


def test_emu_candidates_are_exactly_the_translations_below_the_threshold():
    """dlpd_search_candidates vs the masked score volume of the same call (Docker.update_top feed, Docker.py:86-105)."""
    L, C0, C1, H = 8, 4, 6, 5
    rec, lig, rf, lf = _case(L, C0, C1, seed=31)
    filt = O.FilterWeights.random(C0 + C1, H, seed=2)
    Rs = R20[[11, 500, 1400]]
    tau = float(torch.quantile(O.correlate(rf, lf[None, None]).flatten(), 0.7))
    plan = emu.EmuPlan(L, C0, C1, H, max_rot=2)
    try:
        plan.set_receptor(rec[0][0], rec[1][0], rf[0, 0])
        plan.set_ligand(lig[0][0], lig[1][0], lf)
        plan.set_filter(filt.W1, filt.b1, filt.W2, filt.b2)
        best, scores = plan.search(Rs, tau, want_scores=True)
        thr = float(torch.kthvalue(scores.flatten(), 40).values)
        best2, cs, crot, clin, total = plan.search_candidates(Rs, tau, thr, cap=4096)
        _, _, _, _, total_small = plan.search_candidates(Rs, tau, thr, cap=8)
    finally:
        plan.close()
    assert torch.equal(best, best2)
    flat = scores.reshape(3, -1)
    want = torch.nonzero(flat <= thr)
    assert total == want.shape[0] == total_small and total >= 40
    got = sorted(zip(crot.tolist(), clin.tolist(), cs.tolist()))
    assert [(r, l) for r, l, _ in got] == [tuple(w) for w in want.tolist()]
    assert all(s == float(flat[r, l]) for r, l, s in got)


def test_emu_search_prerotated_equals_search_of_rotated_volumes():
    """dockE3 entry point: volumes rotated by the caller (here: with the oracle's VolumeRotation) must score exactly
    like the fused gather of the same volumes, up to the interpolation rounding of the two rotation codes."""
    L, C0, C1, H = 8, 4, 6, 5
    rec, lig, rf, lf = _case(L, C0, C1, seed=41)
    filt = O.FilterWeights.random(C0 + C1, H, seed=7)
    Rs = R20[[21, 900, 1700]]
    rot = lambda v: torch.stack([O.rotate_volume(v.double(), Rs[i][None])[0] for i in range(3)]).float()
    v0, v1, cl = rot(lig[0]), rot(lig[1]), rot(lf[None, None])[:, 0]
    tau = 4.0
    plan = emu.EmuPlan(L, C0, C1, H, max_rot=2)
    try:
        plan.set_receptor(rec[0][0], rec[1][0], rf[0, 0])
        plan.set_filter(filt.W1, filt.b1, filt.W2, filt.b2)
        best_pre, sc_pre = plan.search_prerotated(v0, v1, cl, tau, want_scores=True)
        plan.set_ligand(lig[0][0], lig[1][0], lf)
        best, sc = plan.search(Rs, tau, want_scores=True)
    finally:
        plan.close()
    ok = (sc_pre != 0) == (sc != 0)                         # threshold-straddling voxels may flip
    assert ok.float().mean() > 0.999
    assert (sc_pre - sc)[ok].abs().max() <= 1e-4 * sc.abs().max()
    d = lambda t: t.double()
    spectra = O.ReceptorSpectra([d(r) for r in rec], d(rf))
    for i in range(3):
        V = O.score_rotation(spectra, [d(v0[i])[None], d(v1[i])[None]], d(cl[i]), torch.eye(3, dtype=torch.float64),
                             filt.to(torch.float64), tau, ligand_forbidden_is_rotated=True)
        s, x, y, z = O.best_pose(V)
        assert [int(v) for v in best_pre[i, 1:]] == [x, y, z]
        assert abs(float(best_pre[i, 0]) - s) <= 1e-4 * max(abs(s), 1e-3 * float(V.abs().max()))


def test_emu_project_clash_matches_oracle():
    """dlpd_project_clash vs the oracle's restatement of Docker.py:221-224 (atoms partly outside the box included)."""
    g = torch.Generator().manual_seed(4)
    L, res = 16, 1.25
    coords = (torch.rand(200, 3, generator=g, dtype=torch.float64) - 0.5) * 22.0       # box is 20 A: some atoms poke out
    Rs = R20[[0, 77, 1500]]
    got = emu.project_clash(coords, Rs, L, res)
    for i in range(3):
        want = O.ligand_clash_volume(coords, Rs[i], L, res)
        assert (got[i] - want).abs().max() <= 2e-6 * max(1.0, float(want.abs().max()))


@pytest.fixture
def pipelined_epilogue():
    emu.set_option(emu.OPT_EPILOGUE_KIND, emu.EPILOGUE_PIPELINED)
    yield
    emu.set_option(emu.OPT_EPILOGUE_KIND, emu.EPILOGUE_AUTO)


@pytest.mark.parametrize("L,C0,C1,H,n", [(8, 4, 6, 5, 3), (20, 4, 6, 5, 3), (16, 16, 32, 24, 2)])
def test_emu_pipelined_epilogue_matches_oracle(L, C0, C1, H, n, pipelined_epilogue):
    test_emu_search_matches_oracle(L, C0, C1, H, n)


def test_emu_pipelined_epilogue_candidates(pipelined_epilogue):
    test_emu_candidates_are_exactly_the_translations_below_the_threshold()


@pytest.fixture
def duo_epilogue():
    emu.set_option(emu.OPT_EPILOGUE_KIND, emu.EPILOGUE_DUO)
    yield
    emu.set_option(emu.OPT_EPILOGUE_KIND, emu.EPILOGUE_AUTO)


@pytest.mark.parametrize("L,C0,C1,H,n", [(8, 4, 6, 5, 3), (20, 4, 6, 5, 3), (16, 16, 32, 24, 2)])
def test_emu_duo_epilogue_matches_oracle(L, C0, C1, H, n, duo_epilogue):
    test_emu_search_matches_oracle(L, C0, C1, H, n)


def test_emu_duo_epilogue_candidates(duo_epilogue):
    test_emu_candidates_are_exactly_the_translations_below_the_threshold()


# ---- fused epilogue (butterfly task of half-tile i interleaved with the filter of half-tile i - 1) -------------------
@pytest.fixture
def fused_epilogue():
    emu.set_option(emu.OPT_EPILOGUE_KIND, emu.EPILOGUE_FUSED)
    yield
    emu.set_option(emu.OPT_EPILOGUE_KIND, emu.EPILOGUE_AUTO)


@pytest.mark.parametrize("L,C0,C1,H,n", [(8, 4, 6, 5, 3), (20, 4, 6, 5, 3), (16, 16, 32, 24, 2)])
def test_emu_fused_epilogue_matches_oracle(L, C0, C1, H, n, fused_epilogue):
    test_emu_search_matches_oracle(L, C0, C1, H, n)


def test_emu_fused_epilogue_candidates(fused_epilogue):
    test_emu_candidates_are_exactly_the_translations_below_the_threshold()


def test_emu_correlate_80_team_plane_matches_oracle():
    """L = 80 is the one size whose plane kernel runs the TEAM path (5 teams, per-team barriers, inverse-x teams owning two
    interleaved column blocks): one channel through the standalone correlation checks its index maps on the CPU."""
    g = torch.Generator().manual_seed(80)
    v1 = torch.randn(1, 1, 80, 80, 80, generator=g)
    v2 = torch.randn(1, 1, 80, 80, 80, generator=g)
    got = emu.volume_correlate(v1, v2)
    want = O.correlate(v1.double(), v2.double())
    assert (got.double() - want).abs().max() <= 2e-5 * want.abs().max()


def test_emu_search_80_one_rotation_matches_oracle():
    """The BASELINE shape (L = 80, 16 + 32 channels) through the emulated kernels for ONE rotation: covers the persistent team
    path of the plane kernel with the tiled pair stores, the 2x2-tile epilogue and HiddenL1 at full size on the CPU."""
    from deeplocalproteindocking_b200 import synthetic
    L, C0, C1, H = 80, 16, 32, 24
    rec, rec_clash = synthetic.learned_volumes(L, (C0, C1), seed=1)
    lig, lig_clash = synthetic.learned_volumes(L, (C0, C1), seed=2)
    filt = O.FilterWeights.random(C0 + C1, seed=3)
    Rs = R20[[316]]
    plan = emu.EmuPlan(L, C0, C1, H, max_rot=2)
    try:
        plan.set_receptor(rec[0], rec[1], rec_clash)
        plan.set_ligand(lig[0], lig[1], lig_clash)
        plan.set_filter(filt.W1, filt.b1, filt.W2, filt.b2)
        best, _ = plan.search(Rs, 30.0)
    finally:
        plan.close()
    want, _ = O.dock([v[None].double() for v in rec], rec_clash[None, None].double(), [v[None].double() for v in lig],
                     lig_clash.double(), Rs, filt.to(torch.float64), 30.0, K=0)
    assert torch.equal(best[:, 1:].double(), want[:, 1:])
    assert torch.allclose(best[:, 0].double(), want[:, 0], rtol=1e-4, atol=1e-6)


# ---- round 2: in-library clash re-projection from atoms, typed projector, per-plan options ---------------------------
def test_emu_search_atoms_equals_explicit_clash_volumes_and_oracle():
    """dlpd_search_atoms (float64 matrices in; fp32 cast + re-projection of the rotated atoms chunk by chunk inside the
    library, Docker.py:214-224) == dlpd_search fed with dlpd_project_clash volumes == the oracle; host-buffer entry too."""
    L, C0, C1, H, res = 8, 4, 6, 5, 1.25
    rec, lig, rf, _ = _case(L, C0, C1, seed=31)
    g = torch.Generator().manual_seed(32)
    atoms = (torch.rand(40, 3, generator=g, dtype=torch.float64) - 0.5) * 6.0
    filt = O.FilterWeights.random(C0 + C1, H, seed=33)
    Rs = R20[[2, 222, 1222]]
    tau = 1.0
    plan = emu.EmuPlan(L, C0, C1, H, max_rot=2)
    try:
        plan.set_receptor(rec[0][0], rec[1][0], rf[0, 0])
        plan.set_ligand(lig[0][0], lig[1][0], None)
        plan.set_filter(filt.W1, filt.b1, filt.W2, filt.b2)
        with pytest.raises(Exception, match="dlpd_set_ligand_atoms"):
            plan.search_atoms(Rs, tau)
        plan.set_ligand_atoms(atoms, res)
        best, scores = plan.search_atoms(Rs, tau, want_scores=True)
        clash = emu.project_clash(atoms, Rs, L, res)
        best2, _ = plan.search(Rs.float(), tau, clash_rot=clash)
        best3 = plan.search_host_atoms(Rs, tau)
    finally:
        plan.close()
    assert torch.equal(best, best2) and torch.equal(best, best3)
    spectra = O.ReceptorSpectra([r.double() for r in rec], rf.double())
    for i in range(3):
        c = O.ligand_clash_volume(atoms, Rs[i], L, res).double()
        V = O.score_rotation(spectra, [l.double() for l in lig], c, Rs[i], filt.to(torch.float64), tau, ligand_forbidden_is_rotated=True)
        s, x, y, z = O.best_pose(V)
        assert [int(v) for v in best[i, 1:]] == [x, y, z]
        assert abs(float(best[i, 0]) - s) <= 1e-4 * max(abs(s), 1e-3 * float(V.abs().max()))


def test_emu_project_typed_matches_torch_path():
    from deeplocalproteindocking_b200 import TypedCoords2Volume
    from deeplocalproteindocking_b200.docker import Atoms, typed_layout
    L, res = 10, 1.25
    g = torch.Generator().manual_seed(5)
    coords = (torch.rand(60, 3, generator=g, dtype=torch.float64) - 0.2) * 14.0
    types = torch.randint(0, 11, (60,), generator=g)
    c, nat, off = typed_layout(Atoms(coords.float(), types), "cpu")
    want = TypedCoords2Volume(L, res)(c, nat, off)
    got = emu.project_typed(c, nat, off, L, res)
    assert (got - want).abs().max() <= 2e-6 * max(1.0, float(want.abs().max()))


def test_emu_options_belong_to_the_plan():
    """dlpd_set_option only seeds new plans; dlpd_plan_set_option changes one plan; a plane-mode switch after set_receptor is an
    error (the spectra are laid out for the other kernel), not a silent corruption."""
    L, C0, C1, H = 16, 4, 6, 5
    rec, lig, rf, lf = _case(L, C0, C1, seed=41)
    filt = O.FilterWeights.random(C0 + C1, H, seed=42)
    Rs = R20[[7, 700]]
    a = emu.EmuPlan(L, C0, C1, H)
    emu.set_option(emu.OPT_PLANE_MODE, emu.PLANE_CLUSTER)        # default for plans created from now on
    b = emu.EmuPlan(L, C0, C1, H)
    emu.set_option(emu.OPT_PLANE_MODE, emu.PLANE_AUTO)
    try:
        out = []
        for p in (a, b):
            p.set_receptor(rec[0][0], rec[1][0], rf[0, 0]); p.set_ligand(lig[0][0], lig[1][0], lf)
            p.set_filter(filt.W1, filt.b1, filt.W2, filt.b2)
            out.append(p.search(Rs, 3.0)[0])
        assert torch.equal(out[0][:, 1:], out[1][:, 1:]) and torch.allclose(out[0][:, 0], out[1][:, 0], rtol=2e-5, atol=1e-6)
        a.set_option(emu.OPT_PLANE_MODE, emu.PLANE_CLUSTER)
        with pytest.raises(Exception, match="plane mode"):
            a.search(Rs, 3.0)
        a.set_option(emu.OPT_PLANE_MODE, emu.PLANE_SINGLE)
        assert torch.equal(a.search(Rs, 3.0)[0], out[0])
    finally:
        a.close(); b.close()


def test_emu_wide_model_16_matches_oracle():
    """Kernel bodies at the wide model (32 + 64 channels, H = 48) through the emulation."""
    test_emu_search_matches_oracle(16, 32, 64, 48, 2)
