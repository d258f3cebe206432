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


def test_tpl_module_surface_is_import_compatible(tmp_path):
    """deeplocalproteindocking_b200.tpl.install() makes the reference's own import lines resolve (Docker.py:8-10,
    VolumeViz.py:10) and the FullAtomModel stand-ins follow the call-site conventions (r' = R r, real atoms only)."""
    import importlib
    import sys
    from deeplocalproteindocking_b200 import tpl
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "_Volume" or k.startswith("TorchProteinLibrary")}
    try:
        tpl.install()
        from TorchProteinLibrary.FullAtomModel import CoordsTransform                                  # Docker.py:8
        from TorchProteinLibrary.Volume import TypedCoords2Volume, VolumeConvolution, VolumeRotation    # Docker.py:9
        from TorchProteinLibrary.FullAtomModel import (PDB2CoordsUnordered, Coords2TypedCoords, CoordsTranslate,   # Docker.py:10
                                                       getBBox, getRandomRotation)
        import _Volume                                                                                   # VolumeViz.py:10
        assert VolumeRotation is importlib.import_module("deeplocalproteindocking_b200").VolumeRotation
        assert VolumeConvolution(clip=5.0) is not None and TypedCoords2Volume(8, 1.0) is not None
        g = torch.Generator().manual_seed(3)
        coords = torch.zeros(2, 3 * 5, dtype=torch.float64)
        coords[0, :12] = torch.randn(12, generator=g, dtype=torch.float64)      # 4 real atoms + 1 padding atom
        coords[1, :9] = torch.randn(9, generator=g, dtype=torch.float64)        # 3 real atoms + 2 padding atoms
        nat = torch.tensor([4, 3], dtype=torch.int32)
        R = getRandomRotation(2)
        assert R.dtype == torch.float64 and torch.allclose(R @ R.transpose(1, 2), torch.eye(3, dtype=torch.float64).expand(2, 3, 3), atol=1e-12)
        assert torch.allclose(torch.linalg.det(R), torch.ones(2, dtype=torch.float64))
        rot = CoordsTransform.CoordsRotate()(coords, R, nat)
        assert torch.allclose(rot[0, :3], R[0] @ coords[0, :3]) and torch.equal(rot[0, 12:], coords[0, 12:])
        T = torch.tensor([[1.0, 2.0, 3.0], [-1.0, 0.5, 0.0]], dtype=torch.float64)
        tr = CoordsTranslate()(coords, T, nat)
        assert torch.allclose(tr[1, 3:6], coords[1, 3:6] + T[1]) and torch.equal(tr[1, 9:], coords[1, 9:])     # padding untouched
        a, b = getBBox(coords, nat)
        xyz0 = coords[0, :12].reshape(4, 3)
        assert torch.equal(a[0], xyz0.min(0).values) and torch.equal(b[0], xyz0.max(0).values)
        with pytest.raises(NotImplementedError):
            PDB2CoordsUnordered()(["x.pdb"])                               # real PDB parsing is outside the search path: loud
        # synthetic proteins by name: same return signature as TorchProteinLibrary's reader / typer (Docker.py:51-52)
        from deeplocalproteindocking_b200 import tpl as T_
        T_.register_synthetic("toy", [[0.0, 0.0, 0.0], [1.0, 2.0, 3.0], [4.0, 5.0, 6.0]], [7, 0, 7])
        c, chains, resnames, resnums, atomnames, na = PDB2CoordsUnordered()(["toy", "synthetic:5:3.0:9"])
        assert c.dtype == torch.float64 and tuple(c.shape) == (2, 15) and na.tolist() == [3, 5] and na.dtype == torch.int32
        assert tuple(resnames.shape) == (2, 5, 4) and tuple(atomnames.shape) == (2, 5, 4) and tuple(chains.shape) == (2, 5)
        tc, nat_t, off_t = Coords2TypedCoords()(c, resnames, atomnames, na)
        assert nat_t.dtype == torch.int32 and tuple(nat_t.shape) == (2, 11) and nat_t.sum(1).tolist() == [3, 5]
        assert nat_t[0].tolist() == [1, 0, 0, 0, 0, 0, 0, 2, 0, 0, 0] and off_t[0].tolist() == [0, 1, 1, 1, 1, 1, 1, 1, 3, 3, 3]
        assert tc[0, :9].tolist() == [1.0, 2.0, 3.0, 0.0, 0.0, 0.0, 4.0, 5.0, 6.0]          # stable sort by type
        out = tmp_path / "v.xplor"
        _Volume.Volume2Xplor(torch.arange(27, dtype=torch.float32).reshape(3, 3, 3), str(out), 1.25)
        text = out.read_text().splitlines()
        assert "ZYX" in text and text[-2].strip() == "-9999"
    finally:
        for k in [k for k in sys.modules if k == "_Volume" or k.startswith("TorchProteinLibrary")]:
            del sys.modules[k]
        for k, m in saved.items():
            if m is not None:
                sys.modules[k] = m


def test_reference_docker_module_imports_against_the_stand_ins(tmp_path):
    """With tpl.install() the reference's UNMODIFIED src/Docker/Docker.py imports and binds our operators (authoring
    container only: the reference tree is absent on the GPU box)."""
    import importlib
    import sys
    from oracle import ref_shims
    if not ref_shims.reference_available():
        pytest.skip("reference tree not present")
    from deeplocalproteindocking_b200 import tpl
    from deeplocalproteindocking_b200.rotations import write_eul, soi_euler_angles
    (tmp_path / "data").mkdir()
    write_eul(str(tmp_path / "data" / "oim20.eul"), soi_euler_angles(20))
    saved = dict(sys.modules)
    try:
        ref_shims.install(str(tmp_path))                 # fake `src` package + se3cnn stubs (test infrastructure)
        tpl.install()                                    # the product's TorchProteinLibrary stand-ins replace the oracle's
        for k in [k for k in sys.modules if k.startswith("src.Docker") or k.startswith("src.Models")]:
            del sys.modules[k]
        mod = importlib.import_module("src.Docker.Docker")
        assert mod.VolumeRotation.__module__ == "deeplocalproteindocking_b200.volume"
        assert mod.VolumeConvolution.__module__ == "deeplocalproteindocking_b200.volume"
        assert mod.getBBox.__module__ == "deeplocalproteindocking_b200.tpl"
        assert mod.CoordsTransform.CoordsRotate.__module__ == "deeplocalproteindocking_b200.tpl"
        models = importlib.import_module("src.Models.DockingModels")
        assert models.VolumeConvolution.__module__ == "deeplocalproteindocking_b200.volume"
        # the CPU half of the reference's dockSE3 set-up, UNMODIFIED, on the synthetic stand-ins (Docker.py:18-61,190-201):
        # constructor, load_batch (reader -> typer -> bbox centring), randomised receptor rotation, box-centre translation
        from deeplocalproteindocking_b200 import synthetic
        torch.manual_seed(0)
        dk = mod.Docker(docking_model=None, angle_inc=20, box_size=20, resolution=1.25, max_conf=3, randomize_rot=True)
        assert dk.rot.R.shape == (1854, 3, 3)
        lig = synthetic.random_protein(40, 6.0, seed=5)
        tpl.register_synthetic("lig", lig.coords, lig.types)
        coords, nat, off, T, natoms = dk.load_batch(["lig", "lig"], bbox_center=False)
        order = torch.argsort(lig.types, stable=True)
        assert torch.allclose(coords[0].reshape(-1, 3), lig.centred().coords[order], atol=1e-12)
        assert nat[0].tolist() == torch.bincount(lig.types, minlength=11).tolist() and natoms.tolist() == [40, 40]
        moved = dk.translate(dk.rotate(coords, dk.randR.repeat(2, 1, 1), natoms), dk.box_center.squeeze().repeat(2, 1), natoms)
        want = lig.centred().coords[order] @ dk.randR[0].t() + 12.5
        assert torch.allclose(moved[1].reshape(-1, 3), want, atol=1e-12)
        dk.cleanup()
    finally:
        for k in [k for k in sys.modules if k not in saved]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_pose_file_round_trip_and_application(tmp_path):
    """Docker.write_conformations -> poses.parse_output gives back rotation, translation and score of every pose; the
    reference reader's int() truncation is available behind a flag; transform_ligand = rotate then translate."""
    from deeplocalproteindocking_b200 import Docker, poses
    dk = Docker(docking_model=None, angle_inc=20, box_size=80, resolution=1.25, max_conf=5, device="cpu")
    dk.top_list = [(7, 3, 157, 80, -2.5), (1853, 0, 79, 159, -1.25), (100, 81, 1, 2, 0.0)]
    path = str(tmp_path / "target.dat")
    assert dk.new_log(path, rewrite=True)
    dk.write_conformations()
    dk.cleanup()
    got = poses.parse_output(path)
    assert poses.parse_output(str(tmp_path / "missing.dat")) is None
    assert len(got) == 3
    for entry, (R, t, score) in zip(dk.top_list, got):
        r_want, t_want, s_want = dk.conformation(entry)
        assert torch.allclose(R[0], r_want, atol=1e-6) and torch.allclose(t[0], t_want, atol=1e-6) and abs(score - s_want) < 1e-6
    # circular index -> signed shift times the resolution (Docker.py:112-122)
    assert torch.allclose(got[0][1][0], torch.tensor([3 * 1.25, -3 * 1.25, -80 * 1.25], dtype=torch.float64))
    trunc = poses.parse_output(path, truncate_translation=True)
    assert torch.equal(trunc[0][1][0], torch.tensor([3.0, -3.0, -100.0], dtype=torch.float64))          # int(3.75), int(-3.75), int(-100.0)
    coords = torch.tensor([[1.0, 0.0, 0.0, 0.0, 2.0, 0.0, 9.0, 9.0, 9.0]], dtype=torch.float64)          # 2 real atoms + padding
    nat = torch.tensor([2], dtype=torch.int32)
    centred = poses.centre_on_bbox(coords, nat)
    assert torch.allclose(centred[0, :6], torch.tensor([0.5, -1.0, 0.0, -0.5, 1.0, 0.0], dtype=torch.float64))
    out = poses.transform_ligand(centred, nat, got[1])
    R, t, _ = got[1]
    assert torch.allclose(out[0, :3], R[0] @ centred[0, :3] + t[0]) and torch.equal(out[0, 6:], coords[0, 6:])


def test_candidate_merge_equals_append_and_stable_sort():
    """`Docker._merge_candidates` (global top-K from the epilogue's candidate lists) must give what the reference gives: append
    every rotation's extractions in rotation order, stable-sort by score, keep K (Docker.py:100-105) -- including for lists
    much longer than K, which are cut with tensor sorts before the Python merge, and with tied scores."""
    import random
    import types
    random.seed(3)
    N, K = 12, 25
    holder = types.SimpleNamespace(top_list=[], max_conf=K)
    reference = []
    sel = torch.tensor([40, 41, 42, 43, 44, 45, 46, 47])
    for batch in range(3):
        m = [400, 7, 60][batch]                                  # one long list (cut by tensor sorts), one short, one medium
        rows = [(random.choice([-1.0, -2.0, -3.5, -0.25, -7.0 + batch]), random.randrange(8), random.randrange(N ** 3)) for _ in range(m)]
        rows = list({(j, lin): (s, j, lin) for s, j, lin in rows}.values())     # a (rotation, voxel) pair occurs once
        random.shuffle(rows)                                     # the device appends in arbitrary order
        cand = torch.tensor(rows, dtype=torch.float64)
        dk.Docker._merge_candidates(holder, cand, sel + 8 * batch, N)
        # reference order: rotations in batch order, each rotation's voxels in row-major order, then ONE stable sort by score
        for s, j, lin in sorted(rows, key=lambda t: (t[1], t[2])):
            reference.append((int(sel[j]) + 8 * batch, lin // (N * N), (lin // N) % N, lin % N, s))
        reference.sort(key=lambda t: t[4])
        reference = reference[:K]
        assert holder.top_list == reference
