"""Generate `tests/golden/*.pt` by RUNNING THE REFERENCE'S OWN CODE (authoring container only).

TEST INFRASTRUCTURE.  `/root/reference` does not exist on the GPU box, so the outputs are committed as
small fixtures and the tests compare the oracle (and, on the GPU, the CUDA path) against them.

    python -m oracle.make_golden

What is executed from the reference, unmodified (through `oracle/ref_shims.py`):
  rotations.pt      `Rotations(angle_inc)` for 20/15/12/10 (Rotations.py:34-66) -> sampled R matrices
  correlation.pt    `MultiplyVolumes.multiply` (MultiplyVolumes.py:13-47) at 40 translations
  global_model.pt   `SimpleFilter`, `GlobalDockingModel.forward`, `LocalDockingModel.forward`
                    (DockingModels.py:23-37,63-84,102-120) with an identity representation
  update_top.pt     `Docker.update_top` (Docker.py:86-105) on masked score volumes
  conformations.pt  `Docker.write_conformations` (Docker.py:107-133) text output
The TPL operators inside those calls are the oracle's (parity unpinned for them, see ref_shims).
"""
from __future__ import annotations

import io
import os

import torch

from . import docking_oracle as O
from . import ref_shims

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), os.pardir, "tests", "golden")


class _Identity(torch.nn.Module):
    """representation stand-in: input is already the list of feature volumes."""

    def forward(self, x):
        return x


def main():
    Rotations, Models, Docker = ref_shims.install(ref_shims.REFERENCE_ROOT)
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(20261017)

    # ---- rotations ------------------------------------------------------------------------------
    rot = {}
    for inc in (20, 15, 12, 10):
        R = Rotations(angle_inc=inc).R
        idx = torch.arange(0, R.shape[0], 53)
        rot[inc] = {"n": R.shape[0], "idx": idx, "R": R[idx].clone()}
    torch.save(rot, os.path.join(OUT, "rotations.pt"))

    # ---- correlation definition -------------------------------------------------------------------
    L, C = 6, 3
    v1 = torch.randn(C, L, L, L, dtype=torch.float64)
    v2 = torch.randn(C, L, L, L, dtype=torch.float64)
    mv = Models.MultiplyVolumes()
    ds = torch.randint(-(L - 1), L, (40, 3))
    ds[:8] = torch.tensor([[sx, sy, sz] for sx in (-2, 3) for sy in (-1, 2) for sz in (-3, 1)])
    vals = torch.stack([mv.multiply(v1, v2, int(d[0]), int(d[1]), int(d[2]), L) for d in ds])
    torch.save({"v1": v1, "v2": v2, "d": ds, "values": vals}, os.path.join(OUT, "correlation.pt"))

    # ---- GlobalDockingModel / LocalDockingModel / SimpleFilter ------------------------------------------
    L0, C0, C1 = 8, 4, 6
    filt = Models.SimpleFilter([C0, C1]).double()
    rec = [torch.randn(1, C0, L0, L0, L0, dtype=torch.float64), torch.randn(1, C1, L0 // 2, L0 // 2, L0 // 2, dtype=torch.float64)]
    lig = [torch.randn(1, C0, L0, L0, L0, dtype=torch.float64), torch.randn(1, C1, L0 // 2, L0 // 2, L0 // 2, dtype=torch.float64)]
    gmodel = Models.GlobalDockingModel(representation=_Identity(), filter=filt)
    with torch.no_grad():
        Vg = gmodel(rec, lig)                                          # [2L,2L,2L] (batch squeezed by the model's resize)
    lmodel = Models.LocalDockingModel(representation=_Identity(), filter=filt)
    Ts = torch.tensor([[0, 0, 0], [2, 4, -6], [-2, -4, 6], [4, 0, 2], [1, 3, 5], [-6, 6, 0], [7, 7, 7], [-1, 2, 3]], dtype=torch.float64)
    # LocalDockingModel.forward computes representation(receptor): feed the lists through the identity module
    with torch.no_grad():
        yl = torch.stack([_local(lmodel, rec, lig, T) for T in Ts])
    torch.save({"rec": rec, "lig": lig, "W1": filt.fc[0].weight.detach(), "b1": filt.fc[0].bias.detach(),
                "W2": filt.fc[2].weight.detach(), "b2": filt.fc[2].bias.detach(), "V_global": Vg.reshape(2 * L0, 2 * L0, 2 * L0),
                "T": Ts, "y_local": yl}, os.path.join(OUT, "global_model.pt"))

    # ---- Docker.update_top -----------------------------------------------------------------------
    docker = Docker(docking_model=gmodel, angle_inc=20, box_size=4, resolution=1.25, max_conf=12, randomize_rot=False)
    cases = []
    g = torch.Generator().manual_seed(5)
    for kind in ("mixed", "few_negative", "all_positive", "positive_with_zero", "ties"):
        V = torch.randn(8, 8, 8, generator=g)
        if kind == "mixed":
            V = V * (torch.rand(8, 8, 8, generator=g) < 0.7).float()
        elif kind == "few_negative":
            V = V.abs(); V[1, 2, 3] = -1.5; V[7, 0, 0] = -0.25; V[0, 5, 5] = -0.25; V[3, 3, 3] = 0.0
        elif kind == "all_positive":
            V = V.abs() + 0.1
        elif kind == "positive_with_zero":
            V = V.abs() + 0.1; V[4, 4, 1] = 0.0; V[2, 0, 6] = -0.0
        elif kind == "ties":
            V = torch.round(V * 2) / 2
        docker.top_list = []
        V_in = V.clone()
        docker.update_top(V.clone(), 3)
        first = list(docker.top_list)
        V2 = torch.roll(V_in, 1, 0) * 0.9
        docker.update_top(V2.clone(), 5)
        cases.append({"kind": kind, "V": V_in, "V2": V2, "K": 12, "after_first": first, "after_second": list(docker.top_list)})
    torch.save(cases, os.path.join(OUT, "update_top.pt"))

    # ---- Docker.write_conformations ------------------------------------------------------------------
    out = {}
    for rand in (False, True):
        dk = Docker(docking_model=gmodel, angle_inc=20, box_size=80, resolution=1.25, max_conf=6, randomize_rot=rand)
        dk.top_list = [(0, 0, 0, 0, -3.5), (17, 79, 80, 159, -2.25), (1853, 81, 3, 120, -1.0), (400, 160 - 1, 1, 80, 0.0)]
        dk.log = io.StringIO()
        dk.write_conformations()
        out["rand" if rand else "plain"] = {"top_list": dk.top_list, "text": dk.log.getvalue(),
                                            "randR": dk.randR.clone() if rand else None}
        dk.log = None
    torch.save(out, os.path.join(OUT, "conformations.pt"))
    print("golden fixtures written to", os.path.normpath(OUT))


def _local(lmodel, rec, lig, T):
    """LocalDockingModel.forward (DockingModels.py:102-120) with list inputs: `receptor.size(...)` is taken
    on the first level, so wrap the list in a tensor-like shim exposing size()."""

    class _L(list):
        def size(self, i):
            return self[0].size(i)

    return lmodel(_L(rec), _L(lig), T.reshape(1, 3)).reshape(())


if __name__ == "__main__":
    main()
