#!/usr/bin/env python
"""SURVEY.md 8(f)-2 at the reference's real setting: `Docker(max_conf=2000)` (local_test.py:55) on the benchmark's 80^3 problem.

    python tests/sweeps/topk_bench.py [--rotations 2048] [--k 2000] [--rot-batch 32] [--parity-rotations 64]

Reports rotations/s of the full `dockSE3` (top-K list maintained, pose file written), how the list was maintained (bootstrap /
candidate / exact batches, candidate-buffer overflows, threshold probes) and -- on the first `--parity-rotations` rotations --
parity of the top-K list against the oracle's restatement of `Docker.update_top` (Docker.py:86-105): same scores within
1e-4 of the largest |score|, every pose the oracle ranks clearly above the K-th score present, nothing from outside.
One JSON line on stdout."""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rotations", type=int, default=2048)
    ap.add_argument("--k", type=int, default=2000)
    ap.add_argument("--rot-batch", type=int, default=32)
    ap.add_argument("--parity-rotations", type=int, default=64)
    args = ap.parse_args()
    from deeplocalproteindocking_b200 import Docker, GlobalDockingModel
    from oracle import docking_oracle as O

    dev = torch.device("cuda", 0)
    rep, filt, rec, lig = bench.build_model(dev)
    model = GlobalDockingModel(rep, filt.to(dev), threshold_clash=bench.TAU).to(dev).eval()
    out = {"box": 80, "K": args.k, "rot_batch": args.rot_batch, "rotation_set": "oim10 (file order)"}

    def run(n_rot, k):
        dk = Docker(model, angle_inc=10, box_size=80, resolution=1.25, max_conf=k, rot_batch=args.rot_batch, device=dev)
        path = os.path.join(tempfile.mkdtemp(), "t.dat")
        dk.new_log(path)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dk.dockSE3(rec, lig, rotation_subset=torch.arange(n_rot))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        stats = {"rotations": n_rot, "seconds": dt, "rotations_per_s": n_rot / dt, "exact_batches": dk.exact_batches,
                 "bootstrap_batches": dk.bootstrap_batches, "candidate_batches": dk.candidate_batches,
                 "overflow_batches": dk.overflow_batches, "probe_searches": dk.probe_searches,
                 "pose_file_lines": sum(1 for _ in open(path))}
        top = list(dk.top_list)
        dk.close()
        return stats, top, dk

    run(2 * args.rot_batch, args.k)                                   # warm-up (plan creation, cuDNN autotune)
    out["topk"], _, _ = run(args.rotations, args.k)
    out["best_only"], _, _ = run(args.rotations, 1)                   # the same driver without the list (K = 1)

    if args.parity_rotations > 0:
        n = args.parity_rotations
        stats, top, dk = run(n, args.k)
        rec_vols, rec_clash, lig_vols, lig_coords, _ = bench.build_problem(dev)
        cpu = lambda t: t.detach().float().cpu()
        f = O.FilterWeights(cpu(filt.fc[0].weight), cpu(filt.fc[0].bias), cpu(filt.fc[2].weight), cpu(filt.fc[2].bias))
        Rs = dk.rot.R[:n]
        rotated = lambda i: O.ligand_clash_volume(lig_coords.cpu(), Rs[i], 80, 1.25)
        torch.set_num_threads(os.cpu_count() or 1)
        t0 = time.perf_counter()
        _, want = O.dock([cpu(v)[None] for v in rec_vols], cpu(rec_clash)[None, None], [cpu(v)[None] for v in lig_vols], None, Rs, f,
                         bench.TAU, K=args.k + 100, ligand_forbidden_rotated=rotated)
        t_oracle = time.perf_counter() - t0
        K = args.k
        got_s = torch.tensor([t[4] for t in top], dtype=torch.float64)
        want_s = torch.tensor([t[4] for t in want[:K]], dtype=torch.float64)
        tol = 1e-4 * float(want_s.abs().max())
        kth = float(want_s[-1])
        got = set(t[:4] for t in top)
        missing = {t[:4] for t in want if t[4] < kth - tol} - got
        foreign = got - {t[:4] for t in want if t[4] <= kth + tol}
        out["parity"] = {"rotations": n, "K": K, "max_abs_score_err": float((got_s - want_s).abs().max()), "tol_abs": tol,
                         "missing_poses": len(missing), "foreign_poses": len(foreign),
                         "identical_order_prefix": next((i for i, (a, b) in enumerate(zip(top, want)) if a[:4] != b[:4]), K),
                         "oracle_seconds": t_oracle, "how": stats}
        out["parity"]["ok"] = bool(out["parity"]["max_abs_score_err"] <= tol and not missing and not foreign)
    print(json.dumps(out), flush=True)
    if "parity" in out and not out["parity"]["ok"]:
        sys.exit(3)


if __name__ == "__main__":
    main()
