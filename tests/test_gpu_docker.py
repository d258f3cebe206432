"""End-to-end `Docker.dockSE3` on the GPU vs the oracle's restatement of the reference loop (small box)."""
import io

import pytest
import torch

from oracle import docking_oracle as O
from deeplocalproteindocking_b200 import Docker, E3MultiResRepr4x4, GlobalDockingModel, SimpleFilter, synthetic
from deeplocalproteindocking_b200.docker import Atoms

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _model(seed=0):
    torch.manual_seed(seed)
    rep = E3MultiResRepr4x4(multiplier=8)
    filt = SimpleFilter(rep.get_num_outputs())
    return GlobalDockingModel(rep, filt, threshold_clash=5.0).to(DEV)


@pytest.mark.parametrize("clash_source,max_conf", [("project", 1), ("rotate", 1), ("project", 6)])
def test_docksE3_matches_oracle_loop(clash_source, max_conf, tmp_path):
    L, res = 20, 1.25
    model = _model()
    rec, lig = synthetic.config_pair(L, seed=3)
    docker = Docker(model, angle_inc=20, box_size=L, resolution=res, max_conf=max_conf, rot_batch=4,
                    clash_source=clash_source, device=DEV)
    subset = torch.arange(0, 1854, 371)
    assert docker.new_log(str(tmp_path / "t.dat"))
    docker.dockSE3(rec, lig, batch_size=2, rotation_subset=subset)
    docker.cleanup()

    # oracle: same volumes (computed by the same plain-PyTorch representation, moved to CPU), reference loop restated
    with torch.no_grad():
        rc, lc = rec.centred(), lig.centred()
        rv, rf = docker._volumes(Atoms(rc.coords + docker.box_center, rc.types))
        lv, lf = docker._volumes(Atoms(lc.coords + docker.box_center, lc.types))
    cpu = lambda t: t.detach().cpu().double()
    filt = O.FilterWeights(cpu(model.filter.fc[0].weight), cpu(model.filter.fc[0].bias), cpu(model.filter.fc[2].weight),
                           cpu(model.filter.fc[2].bias))
    Rs = docker.rot.R[subset]
    rotated = (lambda i: O.ligand_clash_volume(lc.coords, Rs[i], L, res).double()) if clash_source == "project" else None
    want_best, want_top = O.dock([cpu(v)[None] for v in rv], cpu(rf)[None, None], [cpu(v)[None] for v in lv], cpu(lf), Rs,
                                 filt, 5.0, K=max_conf, ligand_forbidden_rotated=rotated)
    got = docker.best_per_rotation.double()
    assert torch.equal(got[:, 1:], want_best[:, 1:])
    assert torch.allclose(got[:, 0], want_best[:, 0], rtol=1e-4, atol=1e-5)
    want_top = [(int(subset[i]), x, y, z, s) for i, x, y, z, s in want_top]
    assert [t[:4] for t in docker.top_list] == [t[:4] for t in want_top]
    lines = open(tmp_path / "t.dat").read().splitlines()
    assert len(lines) == len(docker.top_list) and all(len(l.split("\t")) == 13 for l in lines)
