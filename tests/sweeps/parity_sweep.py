#!/usr/bin/env python
"""parity of the CUDA path on a SEEDED SUBSET of the benchmark's rotation set, on the benchmark's own inputs.

    python tests/sweeps/parity_sweep.py [--rotations 512] [--seed 512]

N rotations drawn without replacement from oim10 (sorted, so the packed pairs mix neighbours of the file order) go through
`dlpd_search_atoms` on the GPU and through the fp32 oracle (CPU PyTorch `torch.fft`, all host threads) one by one; an index
mismatch is re-examined with the fp64 oracle and only counts if that does not score the two translations within 1e-4 of each
other.  One JSON line; exit 3 on a real mismatch or a score outside 1e-4 relative."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rotations", type=int, default=512)
    ap.add_argument("--seed", type=int, default=512)
    args = ap.parse_args()
    from deeplocalproteindocking_b200 import DockingSearch, rotations
    from oracle import docking_oracle as O

    dev = torch.device("cuda", 0)
    rec_vols, rec_clash, lig_vols, lig_coords, filt = bench.build_problem(dev)
    Rall = rotations.Rotations(10).R
    g = torch.Generator().manual_seed(args.seed)
    idx = torch.sort(torch.randperm(Rall.shape[0], generator=g)[:args.rotations]).values
    eng = DockingSearch(80, (16, 32), 24, max_rot=32, device=dev)
    eng.set_receptor(rec_vols, rec_clash); eng.set_ligand(lig_vols, None); eng.set_filter(filt)
    eng.set_ligand_atoms(lig_coords, 1.25)
    best = eng.search_atoms(Rall[idx], bench.TAU)[0].cpu()
    eng.close()
    gpu_best = torch.zeros(Rall.shape[0], 4)
    gpu_best[idx] = best

    torch.set_num_threads(os.cpu_count() or 1)
    cpu = lambda t: t.detach().float().cpu()
    f = O.FilterWeights(cpu(filt.fc[0].weight), cpu(filt.fc[0].bias), cpu(filt.fc[2].weight), cpu(filt.fc[2].bias))
    spectra = O.ReceptorSpectra([cpu(v)[None] for v in rec_vols], cpu(rec_clash)[None, None])
    ligs = [cpu(v)[None] for v in lig_vols]
    poses, t0 = [], time.perf_counter()
    for i in idx.tolist():
        clash = O.ligand_clash_volume(lig_coords.cpu(), Rall[i], 80, 1.25)
        V = O.score_rotation(spectra, ligs, clash, Rall[i], f, bench.TAU, ligand_forbidden_is_rotated=True)
        poses.append((i, O.best_pose(V), float(V.abs().max())))
    out = bench.parity_block(poses, gpu_best, rec_vols, rec_clash, lig_vols, lig_coords, filt, Rall)
    out["oracle_seconds"] = time.perf_counter() - t0
    out["seed"] = args.seed
    out["distinct_best_translations"] = len({tuple(int(v) for v in best[j, 1:]) for j in range(best.shape[0])})
    out["against"] = "fp32 oracle (CPU torch.fft), fp64 for ties; %d seeded oim10 rotations, bench.build_problem inputs" % len(poses)
    print(json.dumps(out), flush=True)
    if out["index_mismatches"] or out["max_rel_score_err"] > out["tol"]:
        sys.exit(3)


if __name__ == "__main__":
    main()
