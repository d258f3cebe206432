"""Scalar-only SE(3) representation restatement (ProteinRepresentationModels.py:23-76): state-dict layout of the se3cnn
modules it stands in for, shapes, and the symmetry that defines it (radial kernels commute with grid rotations)."""
import torch

from deeplocalproteindocking_b200.se3_scalar import SE3MultiResReprScalar, radial_shell_basis


def test_state_dict_layout_matches_se3cnn_modules():
    m = SE3MultiResReprScalar(multiplier=8)
    sd = m.state_dict()
    want = set()
    for seq in ("sequence_res0", "sequence_res1"):
        for i in (0, 2, 4, 6):                       # SE3Convolution at even positions, ScalarActivation (no parameters) between
            want.add("%s.%d.kernel.weight" % (seq, i))
            want.add("%s.%d.kernel.kernel_0_0" % (seq, i))
    assert set(sd) == want
    assert tuple(sd["sequence_res0.0.kernel.kernel_0_0"].shape) == (3, 1, 1, 5, 5, 5)
    assert sd["sequence_res0.0.kernel.weight"].numel() == 16 * 11 * 3
    assert sd["sequence_res1.0.kernel.weight"].numel() == 32 * 16 * 3
    assert m.get_num_outputs() == [16, 32]


def test_checkpoint_basis_overrides_the_local_one():
    """A checkpoint carries its own basis buffers: after load_state_dict the output depends only on the file's contents."""
    torch.manual_seed(0)
    a, b = SE3MultiResReprScalar(multiplier=2), SE3MultiResReprScalar(multiplier=2)
    sd = a.state_dict()
    for k in sd:
        if k.endswith("kernel_0_0"):
            sd[k] = sd[k] * 1.7 + 0.01               # a "foreign" basis
    a.load_state_dict(sd)
    b.load_state_dict(sd)
    x = torch.randn(1, 11, 12, 12, 12)
    ya, yb = a(x), b(x)
    assert torch.equal(ya[0], yb[0]) and torch.equal(ya[1], yb[1])
    assert tuple(ya[0].shape) == (1, 4, 12, 12, 12) and tuple(ya[1].shape) == (1, 8, 6, 6, 6)


def test_radial_kernels_commute_with_grid_rotations():
    torch.manual_seed(1)
    m = SE3MultiResReprScalar(multiplier=2).double()
    x = torch.randn(1, 11, 10, 10, 10, dtype=torch.float64)
    y = m(x)[0]
    for dims in ((2, 3), (3, 4), (2, 4)):
        yr = m(torch.rot90(x, 1, dims))[0]
        assert torch.allclose(yr, torch.rot90(y, 1, dims), atol=1e-10)
    yf = m(torch.flip(x, (2,)))[0]                   # radial kernels are also mirror symmetric
    assert torch.allclose(yf, torch.flip(y, (2,)), atol=1e-10)


def test_basis_is_radial_and_normalised():
    b = radial_shell_basis(5)
    assert torch.allclose(b.reshape(3, -1).norm(dim=1), torch.ones(3), atol=1e-6)
    assert torch.allclose(b, b.transpose(3, 4)) and torch.allclose(b, b.flip(5))
    assert int(b[0].argmax()) == 62                  # innermost shell peaks at the centre voxel


def test_select_model_builds_both_families_and_checkpoints_round_trip(tmp_path):
    """`select_model` mirrors local_train.py:20-47; `GlobalDockingModel.save` / `.load` (DockingModels.py:55-61) round-trip the
    SE(3) representation, basis buffers included."""
    import pytest
    from deeplocalproteindocking_b200 import select_model
    m = select_model("SE3", "SE3MultiResReprScalar")
    assert m.representation.get_num_outputs() == [16, 32] and m.filter.fc[0].in_features == 48
    assert select_model("E3", "E3MultiResRepr4x4").representation.get_num_outputs() == [16, 32]
    with pytest.raises(Exception, match="Unknown"):
        select_model("SO3", "x")
    m.save(str(tmp_path), epoch=3)
    m2 = select_model("SE3", "SE3MultiResReprScalar")
    m2.load(str(tmp_path), epoch=3)
    x = torch.randn(1, 11, 8, 8, 8)
    with torch.no_grad():
        a, b = m.representation(x), m2.representation(x)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
