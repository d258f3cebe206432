"""CPU tests of the host-side logic: ABI surface, top-K extraction, pose file, resume rule, model containers."""
import ctypes
import os
import re

import pytest
import torch

from oracle import docking_oracle as O
from deeplocalproteindocking_b200 import _lib, docker as dk, models, rotations


def test_abi_library_exports_every_declared_symbol():
    """The in-tree libdlpd.so loads without a GPU and exports exactly what include/dlpd.h declares."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "dlpd.h")).read()
    declared = set(re.findall(r"\b(dlpd_[a-z_]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.get_lib()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.dlpd_version() >= 100
    assert lib.dlpd_search_supported(80, 16, 32, 24) == 1 and lib.dlpd_search_supported(80, 16, 32, 23) == 0
    assert lib.dlpd_op_supported(40) == 1 and lib.dlpd_op_supported(41) == 0


def test_package_has_no_cpu_path():
    from deeplocalproteindocking_b200 import DockingSearch, VolumeConvolution
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_lib.DlpdError):
        DockingSearch(8, (4, 6), 5)
    with pytest.raises(_lib.DlpdError):
        VolumeConvolution()(torch.zeros(1, 1, 8, 8, 8), torch.zeros(1, 1, 8, 8, 8))


def test_top_conformations_equals_reference_loop(golden_dir):
    cases = torch.load(os.path.join(golden_dir, "update_top.pt"), weights_only=False)
    for c in cases:
        got = [(3, *t) for t in dk.top_conformations(c["V"], c["K"])]
        got.sort(key=lambda t: t[4])
        assert got[:c["K"]] == c["after_first"], c["kind"]
    g = torch.Generator().manual_seed(2)
    for trial in range(20):
        V = torch.randn(4, 5, 6, generator=g)
        if trial % 3 == 0:
            V = V * (torch.rand(4, 5, 6, generator=g) < 0.4).float()
        if trial % 3 == 1:
            V = V.abs() + 0.5
        for K in (1, 9, 200):
            assert dk.top_conformations(V, K) == O.update_top_literal(V, K)


class _Stub:
    threshold_clash = 1.0


def test_pose_file_format_and_resume_rule(golden_dir, tmp_path):
    g = torch.load(os.path.join(golden_dir, "conformations.pt"), weights_only=False)
    for key in ("plain", "rand"):
        d = dk.Docker(_Stub(), angle_inc=20, box_size=80, resolution=1.25, max_conf=6, randomize_rot=(key == "rand"),
                      device="cpu")
        if key == "rand":
            d.randR = g[key]["randR"]
        d.top_list = g[key]["top_list"]
        path = str(tmp_path / (key + ".dat"))
        assert d.new_log(path, rewrite=False)
        d.write_conformations()
        d.cleanup()
        assert open(path).read() == g[key]["text"]
        assert d.new_log(path, rewrite=False) is False             # >1 line present: target is skipped (Docker.py:67-77)
        assert d.new_log(path, rewrite=True) is True
        d.cleanup()


def test_filter_and_representation_shapes_and_state_dict_keys():
    rep = models.E3MultiResRepr4x4(multiplier=8)
    assert rep.get_num_outputs() == [16, 32]
    keys = list(rep.state_dict().keys())
    assert keys == ["conv1.0.weight", "conv1.2.weight", "conv1.4.weight", "conv1.6.weight", "conv1.8.weight",
                    "conv2.1.weight", "conv2.3.weight", "conv2.5.weight", "conv2.7.weight"]
    with torch.no_grad():
        v0, v1 = rep(torch.zeros(1, 11, 16, 16, 16))
    assert v0.shape == (1, 16, 16, 16, 16) and v1.shape == (1, 32, 8, 8, 8)
    f = models.SimpleFilter([16, 32])
    assert list(f.state_dict().keys()) == ["fc.0.weight", "fc.0.bias", "fc.2.weight", "fc.2.bias"]
    assert f.fc[0].weight.shape == (24, 48) and f.fc[2].weight.shape == (1, 24)


def test_projection_matches_oracle_and_typed_layout():
    from deeplocalproteindocking_b200 import TypedCoords2Volume, synthetic
    atoms = synthetic.random_protein(200, 8.0, seed=4)
    L, res = 16, 1.25
    moved = dk.Atoms(atoms.coords + L * res / 2, atoms.types)
    coords, nat, off = dk.typed_layout(dk.Atoms(moved.coords.float(), moved.types), "cpu")
    vol = TypedCoords2Volume(L, res)(coords, nat, off)
    want = O.project_atoms(moved.coords.float(), moved.types, 11, L, res)
    assert torch.allclose(vol[0], want, atol=1e-6)
    batched = dk.project_density_batched(moved.coords.float()[None], L, res)
    assert torch.allclose(batched[0], want.sum(0), atol=1e-5)


def test_bench_torch_loop_body_matches_oracle():
    """bench.py's same-GPU comparator (the reference loop body as PyTorch library calls) computes what the oracle computes."""
    import bench
    from oracle import docking_oracle as O
    from deeplocalproteindocking_b200 import rotations, synthetic
    L, C0, C1 = 8, 4, 6
    rec, rec_clash = synthetic.learned_volumes(L, (C0, C1), seed=61)
    lig, lig_clash = synthetic.learned_volumes(L, (C0, C1), seed=62)
    filt = O.FilterWeights.random(C0 + C1, 5, seed=63)
    R = rotations.Rotations(20).R[321]
    tau = 2.0
    clash_rot = O.rotate_volume(lig_clash[None, None].double(), R[None])[0, 0].float()
    spectra = O.ReceptorSpectra([v[None].double() for v in rec], rec_clash[None, None].double())
    V = O.score_rotation(spectra, [v[None].double() for v in lig], clash_rot.double(), R, filt.to(torch.float64), tau,
                         ligand_forbidden_is_rotated=True)
    s, x, y, z = O.best_pose(V)
    for pre in (False, True):
        sp = None
        if pre:
            sp = [torch.fft.rfftn(v[None], s=(2 * v.shape[-1],) * 3, dim=(-3, -2, -1)) for v in rec]
            sp.append(torch.fft.rfftn(rec_clash[None, None], s=(2 * L,) * 3, dim=(-3, -2, -1)))
        best, idx = bench.torch_loop_body(rec, rec_clash, lig, clash_rot, R.float(), filt.W1, filt.b1, filt.W2, filt.b2, tau, sp)
        N = 2 * L
        assert int(idx) == (x * N + y) * N + z
        assert abs(float(best) - s) <= 1e-4 * max(abs(s), 1e-3 * float(V.abs().max()))
