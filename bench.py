#!/usr/bin/env python
"""bench.py -- rotations/s of the exhaustive docking search (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1]): single receptor-ligand pair, synthetic atoms -> Gaussian densities ->
randomly initialised E3MultiResRepr4x4(multiplier=8) -> 16@80^3 + 32@40^3 volumes, 48->24->1 SimpleFilter, the WHOLE
oim10 rotation set (14 868 rotations, file order), fp32, every one of the 160^3 translations scored per rotation.

The timed region is the whole job: the K steps together cover the rotation set exactly once (a step = ceil(14868 / (K N))
rotations per GPU), so the run lasts seconds, clocks are sampled under sustained load, and N > 1 is STRONG scaling: the
set is sharded i mod N, every rank walks its shard, and the region closes with the one exchange of the path -- the
all-gather of the per-rotation best poses -- plus the global best-pose selection.  Per step and GPU: float64 rotation
matrices -> fp32 cast + re-projection of the rotated ligand atoms into clash volumes (Docker.py:214-224) + fused search
(Docker.py:225-236), all inside libdlpd.so (`dlpd_search_atoms`; the `e2e` leg goes through `dlpd_search_host_atoms` with
HOST buffers).  Per-rank set-up (plan, receptor spectra, ligand upload) is timed and reported separately.

Printed JSON line: see DESIGN.md "Measurement".  `--impl reference` times the CPU oracle (the reference correlation as CPU
PyTorch with torch.fft, all host threads) on a bounded sample of the same workload instead.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

L, C0, C1, HID = 80, 16, 32, 24
ANGLE_INC = 10
TAU = 300.0

# BASELINE.json configs (index = position in BASELINE.json "configs", 1-based).  The bench line the driver reads is
# config 2; 3-5 are the multi-GPU / batch / large-box cases, runnable with --config for measurement.
CONFIGS = {
    2: dict(box=80, angle_inc=10, pairs=1, source="atoms",
            text="single pair, 80^3 grid, oim10 rotation set (14868 rotations, file order)"),
    3: dict(box=80, angle_inc=4, pairs=1, source="atoms",
            text="single pair, 80^3 grid, oim04 rotation set (232020 rotations, deterministic surrogate: the table is a "
                 "missing blob in the reference snapshot)"),
    4: dict(box=80, angle_inc=6, pairs=64, source="learned",
            text="batch of 64 synthetic receptor-ligand pairs (multi-channel learned volumes), oim06 rotation set (68760 "
                 "rotations), pairs dealt round-robin to the GPUs, one pair switch (receptor spectra + ligand upload) per step"),
    5: dict(box=128, angle_inc=4, pairs=1, source="learned",
            text="large-complex 128^3 grid, multi-channel learned volumes, oim04 rotation set (232020 rotations, surrogate)"),
}


def algorithmic_bytes_per_rotation(L=None, C0=C0, C1=C1):
    """per channel of edge l (n = 2l, h = n*n*(n/2+1)): 4 l^3 + 24 h bytes; split by kernel as
    FwdZ = 4 l^3 (read ligand), Plane = 16 h (receptor spectrum read + spectrum write), Epilogue (level 0) and
    HiddenL1 (level 1) = 8 h (read back)."""
    L = globals()["L"] if L is None else L

    def h(l):
        n = 2 * l
        return n * n * (n // 2 + 1)
    lv0, lv1 = (C0 + 1), C1
    parts = {
        "fwdz_level0": lv0 * 4 * L ** 3, "fwdz_level1": lv1 * 4 * (L // 2) ** 3,
        "plane_level0": lv0 * 16 * h(L), "plane_level1": lv1 * 16 * h(L // 2),
        "epilogue": lv0 * 8 * h(L), "hidden_level1": lv1 * 8 * h(L // 2),
    }
    parts["total"] = sum(parts.values())
    return parts


def measured_traffic():
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture at the bench's
    launch shape), committed under profiles/ by tools/ncu_traffic.py; None if no capture matches."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def build_model(device):
    """The synthetic pair and the randomly initialised networks of BASELINE.json configs 1-3 (once, outside any timed region):
    returns (representation, filter, receptor Atoms, ligand Atoms), atoms bbox-centred."""
    from deeplocalproteindocking_b200 import E3MultiResRepr4x4, SimpleFilter, synthetic
    from deeplocalproteindocking_b200.docker import Atoms, TypedCoords2Volume, typed_layout
    torch.manual_seed(1234 + 1)
    rep = E3MultiResRepr4x4(multiplier=8).eval()
    filt = SimpleFilter(rep.get_num_outputs())
    rec, lig = synthetic.config_pair(L, seed=1234 + 1)
    rec, lig = rec.centred(), lig.centred()
    centre = L * 1.25 / 2.0
    project = TypedCoords2Volume(L, 1.25)
    # Gain calibration of the randomly initialised network (always on the CPU, so both bench arms build the same model): five
    # bias-free xavier convolutions shrink the signal to ~1e-3, the correlations then vanish against the filter's biases and
    # EVERY rotation's best pose degenerates to the first masked voxel.  Scaling the last layer of each stack (the network is
    # positively homogeneous: no biases, ReLU / max-pool only) to unit output deviation on the receptor keeps it a plain
    # random-init E3MultiResRepr4x4 and gives score volumes with structure and negative minima.
    with torch.no_grad():
        c0, n0, o0 = typed_layout(Atoms((rec.coords + centre).float(), rec.types), "cpu")
        v = rep(project(c0, n0, o0))
        g0, g1 = 1.0 / float(v[0].std()), 1.0 / float(v[1].std())
        rep.conv1[-1].weight.mul_(g0)
        rep.conv2[-1].weight.mul_(g1 / g0)
    return rep.to(device), filt, rec, lig


def build_problem(device):
    """Synthetic pair -> volumes (once, outside any timed region)."""
    from deeplocalproteindocking_b200.docker import Atoms, TypedCoords2Volume, typed_layout
    rep, filt, rec, lig = build_model(device)
    centre = L * 1.25 / 2.0
    project = TypedCoords2Volume(L, 1.25)

    def volumes(atoms):
        coords, nat, off = typed_layout(Atoms((atoms.coords + centre).float(), atoms.types), device)
        with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            dens = project(coords, nat, off)
            vols = rep(dens)
        return [v[0].contiguous() for v in vols], dens.sum(dim=1)[0].contiguous()

    rec_vols, rec_clash = volumes(rec)
    lig_vols, _ = volumes(lig)
    return rec_vols, rec_clash, lig_vols, lig.coords, filt


def clash_volumes(lig_coords_dev, Rb64, device):
    from deeplocalproteindocking_b200.volume import project_clash_rotated
    return project_clash_rotated(lig_coords_dev, Rb64, L, 1.25)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def cpu_baseline(rec_vols, rec_clash, lig_vols, lig_coords, filt, Rs, n_rot, warm=1, retransform_receptor=False, poses=None):
    """The oracle (CPU PyTorch + torch.fft restatement of the reference loop) timed on all host cores over `n_rot` rotations;
    receptor spectra precomputed (kind to the baseline) or, with `retransform_receptor`, recomputed for every rotation like
    the reference's VolumeConvolution does (DockingModels.py:71).  kind = "port"."""
    from oracle import docking_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    cpu = lambda t: t.detach().float().cpu()
    f = O.FilterWeights(cpu(filt.fc[0].weight), cpu(filt.fc[0].bias), cpu(filt.fc[2].weight), cpu(filt.fc[2].bias))
    spectra = O.ReceptorSpectra([cpu(v)[None] for v in rec_vols], cpu(rec_clash)[None, None])
    ligs = [cpu(v)[None] for v in lig_vols]
    lc = lig_coords.cpu()
    times = []
    if poses is not None:
        poses.clear()
    for i in range(warm + n_rot):
        t0 = time.perf_counter()
        if retransform_receptor:
            spectra = O.ReceptorSpectra([cpu(v)[None] for v in rec_vols], cpu(rec_clash)[None, None])
        clash = O.ligand_clash_volume(lc, Rs[i % Rs.shape[0]], L, 1.25)
        V = O.score_rotation(spectra, ligs, clash, Rs[i % Rs.shape[0]], f, TAU, ligand_forbidden_is_rotated=True)
        pose = O.best_pose(V)
        if i >= warm:
            times.append(time.perf_counter() - t0)
        if poses is not None:
            poses.append((i % Rs.shape[0], pose, float(V.abs().max())))      # (rotation index, (score, x, y, z), max |V|)
    return sum(times), len(times)


def parity_block(poses, gpu_best, rec_vols, rec_clash, lig_vols, lig_coords, filt, Rs, tol=1e-4):
    """The bench's own inputs against the oracle (north_star: best translation indices identical unless the reference scores
    are tied within tolerance, scores within 1e-4 relative).  `poses` come from the cpu_baseline leg (fp32 oracle); an index
    mismatch is re-examined with the fp64 oracle and only counts if that does not call the two translations a tie."""
    from oracle import docking_oracle as O
    cpu64 = lambda t: t.detach().double().cpu()
    mism, ties, max_rel = 0, 0, 0.0
    spectra64 = None
    for rot, (s, x, y, z), vmax in poses:
        g = gpu_best[rot]
        gs, gi = float(g[0]), [int(g[1]), int(g[2]), int(g[3])]
        scale = max(abs(s), 1e-3 * vmax)
        if gi == [x, y, z]:
            max_rel = max(max_rel, abs(gs - s) / scale)
            continue
        if spectra64 is None:
            spectra64 = O.ReceptorSpectra([cpu64(v)[None] for v in rec_vols], cpu64(rec_clash)[None, None])
            f64 = O.FilterWeights(cpu64(filt.fc[0].weight), cpu64(filt.fc[0].bias), cpu64(filt.fc[2].weight), cpu64(filt.fc[2].bias))
            ligs64 = [cpu64(v)[None] for v in lig_vols]
        clash = O.ligand_clash_volume(lig_coords.cpu(), Rs[rot], L, 1.25).double()
        V = O.score_rotation(spectra64, ligs64, clash, Rs[rot], f64, TAU, ligand_forbidden_is_rotated=True)
        a, b = float(V[x, y, z]), float(V[gi[0], gi[1], gi[2]])
        # a tie in fp64, or a mask flip: the clash correlation of one of the two translations sits within tol of the threshold
        n = O.correlate(cpu64(rec_clash)[None, None], clash[None, None])[0, 0]
        near_tau = min(abs(float(n[x, y, z]) - TAU), abs(float(n[gi[0], gi[1], gi[2]]) - TAU)) <= tol * float(n.abs().max())
        if abs(a - b) <= tol * max(abs(a), 1e-3 * float(V.abs().max())) or near_tau:
            ties += 1
        else:
            mism += 1
    return {"rotations": len(poses), "index_mismatches": mism, "ties_or_mask_flips_within_tol": ties, "max_rel_score_err": max_rel, "tol": tol,
            "against": "oracle (CPU PyTorch torch.fft restatement, fp32; fp64 for tie adjudication) on the bench's own inputs: "
                       "oim10 rotations %s" % ([p[0] for p in poses][:3] + ["..."] + [poses[-1][0]] if poses else [])}


def fp32_lane_ops_per_pair(box, filter_kind):
    """FP32 lane-operations (one fp32 add / mul / fma of one lane) per packed rotation PAIR, DESIGN.md section 3: FFT
    butterflies 10.1 G at 80^3 (scaled with N^3 log2 N elsewhere) + filter: per translation and rotation 16 x 24 (level-0
    first layer; on the tensor cores with filter_kind = tcgen05, leaving the hi/lo split and the accumulator add)
    + 32 x 24 / 8 (level-1 share) + 24 (second layer)."""
    import math
    n = 2 * box
    fft = 10.1e9 * (n ** 3 * math.log2(n)) / (160 ** 3 * math.log2(160))
    first = (16 * 2 + 24) if filter_kind == "tcgen05" else 16 * 24
    filt = 2.0 * n ** 3 * (first + 32 * 24 / 8 + 24 * 2)
    return fft + filt


def torch_loop_body(rec_vols, rec_clash, lig_vols, lig_clash_rot, R, W1, b1, W2, b2, tau, rec_spectra=None):
    """ONE rotation of the reference's loop body (Docker.py:218-232 + DockingModels.py:63-84) written with PyTorch LIBRARY
    calls on whatever device the tensors live on -- the formulation the reference itself runs (grid_sample rotation, FFT
    correlation per level, nearest up-sampling, cat / transpose / contiguous, two Linear layers, mask, min).  It is the
    same-GPU comparator of SURVEY.md 8(d) (TorchProteinLibrary's compiled operators cannot be built here); never used by
    the product path.  rec_spectra=None re-transforms the receptor on every call like VolumeConvolution does.
    Returns (best score, flat argmin index)."""
    import torch.nn.functional as F
    dev, N = rec_vols[0].device, 2 * rec_vols[0].shape[-1]
    convs = []
    for k, (rv, lv) in enumerate(zip(rec_vols, lig_vols)):
        Lk, n = lv.shape[-1], 2 * lv.shape[-1]
        ax = torch.arange(Lk, device=dev, dtype=torch.float32) - 0.5 * Lk
        P = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), dim=-1)
        S = torch.einsum("xyzb,ba->xyza", P, R) + 0.5 * Lk                      # source voxel of every output voxel
        g = 2.0 * S / (Lk - 1) - 1.0
        grid = torch.stack([g[..., 2], g[..., 1], g[..., 0]], dim=-1)[None]
        rot = F.grid_sample(lv[None], grid, mode="bilinear", padding_mode="zeros", align_corners=True)
        spec = rec_spectra[k] if rec_spectra is not None else torch.fft.rfftn(rv[None], s=(n, n, n), dim=(-3, -2, -1))
        corr = torch.fft.irfftn(spec * torch.conj(torch.fft.rfftn(rot, s=(n, n, n), dim=(-3, -2, -1))), s=(n, n, n), dim=(-3, -2, -1))
        convs.append(corr if n == N else F.interpolate(corr, scale_factor=N // n, mode="nearest"))
    x = torch.cat(convs, dim=1).transpose(1, 2).transpose(2, 3).transpose(3, 4).contiguous()      # DockingModels.py:79-81
    score = F.linear(torch.relu(F.linear(x.reshape(-1, x.shape[-1]), W1, b1)), W2.reshape(1, -1), b2.reshape(-1)).reshape(N, N, N)
    cspec = rec_spectra[-1] if rec_spectra is not None else torch.fft.rfftn(rec_clash[None, None], s=(N, N, N), dim=(-3, -2, -1))
    clash = torch.fft.irfftn(cspec * torch.conj(torch.fft.rfftn(lig_clash_rot[None, None], s=(N, N, N), dim=(-3, -2, -1))),
                             s=(N, N, N), dim=(-3, -2, -1))[0, 0]
    V = (clash < tau).float() * score                                           # Docker.py:226,232
    best, idx = torch.min(V.reshape(-1), dim=0)                                 # first extraction of update_top
    return best, idx


def torch_gpu_baseline(rec_vols, rec_clash, lig_vols, clash_rot, Rs, filt, n_rot, precomputed, device):
    """rotations/s of torch_loop_body on `device`, timed with CUDA events after one warm-up."""
    W1, b1, W2, b2 = (t.detach().float().to(device) for t in (filt.fc[0].weight, filt.fc[0].bias, filt.fc[2].weight, filt.fc[2].bias))
    spectra = None
    if precomputed:
        spectra = [torch.fft.rfftn(v[None], s=(2 * v.shape[-1],) * 3, dim=(-3, -2, -1)) for v in rec_vols]
        spectra.append(torch.fft.rfftn(rec_clash[None, None], s=(2 * L,) * 3, dim=(-3, -2, -1)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        for i in range(n_rot + 1):
            if i == 1:
                e0.record()
            best, idx = torch_loop_body(rec_vols, rec_clash, lig_vols, clash_rot[i % clash_rot.shape[0]], Rs[i].float().to(device),
                                        W1, b1, W2, b2, TAU, spectra)
            float(best)                                                          # the reference reads the minimum back every step
        e1.record()
    torch.cuda.synchronize(device)
    return n_rot / (1e-3 * e0.elapsed_time(e1))


def run_reference(args, rank, world):
    """`--impl reference`: the reference's own CPU formulation of the path (oracle port; TorchProteinLibrary's
    compiled operators are not vendored) on the host cores.  Rank 0 only."""
    if rank != 0:
        return
    from deeplocalproteindocking_b200 import rotations
    rec_vols, rec_clash, lig_vols, lig_coords, filt = build_problem("cpu")
    Rs = rotations.Rotations(ANGLE_INC).R
    per_step = args.ref_rot_per_step
    cpu_baseline(rec_vols, rec_clash, lig_vols, lig_coords, filt, Rs, per_step * args.warmup, warm=0)
    t, n = cpu_baseline(rec_vols, rec_clash, lig_vols, lig_coords, filt, Rs[per_step * args.warmup:], per_step * args.steps, warm=0)
    value = n / t
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": "rotations/sec (all translations scored)", "value": value, "unit": "rotations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(1),
        "step": {"rotations_per_step": per_step, "note": "bounded sample of the same rotation set (file order after the warm-up rotations)"},
        "cpu_baseline": {"value": value, "unit": "rotations/s", "cores": cores, "kind": "port",
                         "sample": "%d rotations of oim10 per step x %d steps, CPU PyTorch torch.fft, receptor spectra precomputed"
                                   % (per_step, args.steps)},
        "e2e": {"value": value, "unit": "rotations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(world, config=2):
    """What is computed -- identical for both arms; how a step samples it is reported separately under "step"."""
    c = CONFIGS[config]
    return {"workload": "%s, %d@%d^3+%d@%d^3 channels + clash channel, 48->24->1 filter, all %d^3 translations scored per "
                        "rotation, fp32" % (c["text"], C0, c["box"], C1, c["box"] // 2, 2 * c["box"]),
            "baseline_config": config,
            "box_size": c["box"], "channels": [C0, C1], "rotation_set": "oim%02d" % c["angle_inc"],
            "l2_policy": "working set per step >> L2 (receptor spectra + per-pair workspace, GBs); no explicit flush"}


def build_learned_pair(box, seed, device):
    """Configs 4/5: "multi-channel learned volumes" generated directly; the clash volume is rotated with the features."""
    from deeplocalproteindocking_b200 import synthetic
    rec, rec_clash = synthetic.learned_volumes(box, (C0, C1), seed=2 * seed, device=device)
    lig, lig_clash = synthetic.learned_volumes(box, (C0, C1), seed=2 * seed + 1, device=device)
    return rec, rec_clash, lig, lig_clash


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rot-batch", type=int, default=0,
                    help="rotations per step per GPU; 0 = the whole rotation set over the K steps (config 2, strong scaling) "
                         "or 64 (configs 3-5 unless --full-set)")
    ap.add_argument("--full-set", action="store_true", help="configs 3 / 5: cover the whole rotation set too (minutes)")
    ap.add_argument("--max-rot", type=int, default=32, help="rotations per internal chunk (workspace size)")
    ap.add_argument("--cpu-rotations", type=int, default=16, help="rotations of the cpu_baseline / parity sample (0 = skip)")
    ap.add_argument("--torch-gpu-rotations", type=int, default=6,
                    help="rotations of the same-GPU PyTorch (cuFFT / grid_sample) comparator sample (0 = skip)")
    ap.add_argument("--ref-rot-per-step", type=int, default=2)
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json config (1-based); 2 is the bench line")
    ap.add_argument("--plane-mode", default="auto", choices=["auto", "single", "cluster"], help="plane-stage kernel (tuning)")
    ap.add_argument("--epilogue-kind", default="auto", choices=["auto", "classic", "pipelined", "duo", "fused", "tc_duo", "pipe_tc"], help="epilogue kernel (tuning; non-classic kinds need an experimental build)")
    ap.add_argument("--filter-kind", default="auto", choices=["auto", "fp32", "tensor", "tcgen05"], help="first filter layer: tcgen05/TMEM (auto), FFMA2 chain, or mma.sync (experimental build)")
    ap.add_argument("--epilogue-tile", type=int, default=0, choices=[0, 1, 2], help="epilogue tile shape (tuning)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    args.steps = max(args.steps, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from deeplocalproteindocking_b200 import DockingSearch, rotations, set_option
    from deeplocalproteindocking_b200 import distributed as D
    set_option("plane_mode", args.plane_mode)
    set_option("epilogue_tile", args.epilogue_tile)
    set_option("epilogue_kind", args.epilogue_kind)
    set_option("filter_kind", args.filter_kind)

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    cfg = CONFIGS[args.config]
    box = cfg["box"]
    global L
    L = box                                                                   # clash_volumes / cpu_baseline read the module constant
    Rall = rotations.Rotations(cfg["angle_inc"], allow_surrogate=True).R     # [N_rot,3,3] float64 CPU
    n_rot_total = int(Rall.shape[0])
    multi_pair = cfg["pairs"] > 1
    atoms_path = cfg["source"] == "atoms"
    full_set = (args.rot_batch == 0 and args.config == 2) or (args.full_set and not multi_pair)
    shard = D.shard_indices(n_rot_total, rank, world) if not multi_pair else torch.arange(n_rot_total)
    if full_set:
        # the K steps cover the shard exactly once; a step is a whole number of the library's internal chunks (max_rot
        # rotations) so that only the last step of a rank ends in a partial chunk
        nb = -(-D.padded_shard_size(n_rot_total, world) // args.steps)
        nb = -(-nb // args.max_rot) * args.max_rot
    else:
        nb = args.rot_batch or 64
    n_steps_total = args.warmup + args.steps

    def step_indices(s):
        """Timed step k = s - warmup: rotations [k nb, (k+1) nb) of this rank's shard (the last one may be short or empty);
        warm-up steps replay the first slices with the same shapes."""
        k = s - args.warmup if s >= args.warmup else s % args.steps
        if full_set:
            return shard[k * nb:(k + 1) * nb]
        return shard[(torch.arange(nb) + s * nb) % shard.numel()]

    # ---------------- set-up, timed separately (once per rank; replicated) ----------------------------
    def timed(fn):
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize(device)
        return out, 1e3 * (time.perf_counter() - t0)

    if atoms_path:
        (rec_vols, rec_clash, lig_vols, lig_coords, filt), t_problem = timed(lambda: build_problem(device))
        pairs = [(rec_vols, rec_clash, lig_vols, None)]
    else:
        from deeplocalproteindocking_b200 import SimpleFilter
        torch.manual_seed(1234 + args.config)
        filt = SimpleFilter([C0, C1])
        my_pairs = [p for p in range(cfg["pairs"]) if p % world == rank] or [rank % cfg["pairs"]]
        my_pairs = my_pairs[:max(1, min(len(my_pairs), n_steps_total))]
        pairs, t_problem = timed(lambda: [build_learned_pair(box, 1000 * args.config + p, device) for p in my_pairs])
        lig_coords = None
    eng, t_plan = timed(lambda: DockingSearch(box, (C0, C1), HID, max_rot=args.max_rot, device=device))
    eng.set_filter(filt)
    tau = TAU if atoms_path else 30.0
    _, t_rec = timed(lambda: eng.set_receptor(pairs[0][0], pairs[0][1]))
    _, t_lig = timed(lambda: (eng.set_ligand(pairs[0][2], pairs[0][3]),
                              eng.set_ligand_atoms(lig_coords, 1.25) if atoms_path else None))
    setup_ms = {"synthetic_problem": t_problem, "plan_create": t_plan, "set_receptor": t_rec, "set_ligand": t_lig,
                "note": "per rank, before the timed region (receptor spectra are recomputed on every rank, SURVEY.md 8e)"}

    def load_pair(k):
        rv, rc, lv, lc = pairs[k % len(pairs)]
        eng.set_receptor(rv, rc)
        eng.set_ligand(lv, lc)

    R_dev = [Rall[step_indices(s)].to(device) for s in range(n_steps_total)]             # resident inputs, float64 like r_batch_cpu
    R_pin = [Rall[step_indices(s)].clone().pin_memory() for s in range(n_steps_total)]   # host inputs of the e2e leg (float64)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(s):
        if R_dev[s].shape[0] == 0:
            return torch.empty(0, 4, device=device)
        if multi_pair:
            load_pair(s)
        if atoms_path:
            return eng.search_atoms(R_dev[s], tau)[0]
        return eng.search(R_dev[s].to(torch.float32), tau)[0]

    def host_step(s):
        """End to end through the C ABI with HOST buffers (dlpd_search_host_atoms / dlpd_search_host): H2D of the step's
        rotation matrices from pinned memory, clash re-projection + fused search, D2H of the step's best poses, stream sync."""
        if R_pin[s].shape[0] == 0:
            return torch.empty(0, 4)
        if multi_pair:
            load_pair(s)
        if atoms_path:
            return eng.search_host_atoms(R_pin[s], tau)
        return eng.search_host(R_pin[s].float(), tau)

    def finish(bests):
        """The one exchange of the path + what dockSE3 does with it: gather the per-rotation best poses of all ranks
        (rotation order restored), pick the global best pose (Docker.py:100-105 with K = 1)."""
        local = torch.cat(bests) if bests else torch.empty(0, 4, device=device)
        if not full_set:
            if world > 1:
                out = torch.empty(world * local.shape[0], 4, device=device)
                dist.all_gather_into_tensor(out, local)
                return out
            return local
        allb = D.gather_best(local, n_rot_total, rank, world)
        top = torch.argmin(allb[:, 0])
        return allb, top

    # ---------------- device-resident leg (value, roofline) -------------------------------------------------
    for s in range(args.warmup):
        device_step(s)
    if full_set:                                     # the closing collective once outside the region too (NCCL connects lazily)
        D.gather_best(torch.zeros(shard.numel(), 4, device=device), n_rot_total, rank, world)
    barrier()
    eng.profile(True)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    bests = [device_step(s) for s in range(args.warmup, n_steps_total)]
    result = finish(bests)
    ev1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    launches = eng.launch_count() - launches0
    prof = eng.profile_read()
    eng.profile(False)
    n_done = n_rot_total if full_set else world * nb * args.steps              # rotations the whole job scored
    value = n_done / (ms_total * 1e-3)
    gpu_all = result[0].cpu() if full_set else None

    # ---------------- end-to-end leg (host buffers) -----------------------------------------------------------
    for s in range(min(args.warmup, 2)):
        host_step(s)
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    hb = [host_step(s) for s in range(args.warmup, n_steps_total)]
    if full_set:
        D.gather_best(torch.cat(hb).to(device), n_rot_total, rank, world)
    ev1.record()
    barrier()
    e2e_ms = torch.tensor([max(ev0.elapsed_time(ev1), 1e3 * (time.perf_counter() - t0))], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = n_done / (float(e2e_ms.item()) * 1e-3)
    e2e_identical = bool(torch.equal(torch.cat(hb), torch.cat(bests).cpu())) if full_set else None

    rc = 0
    if rank == 0:
        peak, peak_src = measured_peaks()
        alg = algorithmic_bytes_per_rotation(box)
        traffic = measured_traffic()
        rot_rank = (shard.numel() if full_set else nb * args.steps)                 # rotations rank 0 scored in the region
        per_kernel = {}
        for k, (tms, n) in prof.items():
            if n == 0:
                continue
            rot_per_launch = rot_rank / n
            gbs = alg[k] * rot_per_launch / (tms / n * 1e-3) / 1e9
            per_kernel[k] = {"ms_per_launch": tms / n, "launches": n, "share_of_step": tms / ms_total,
                             "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / peak}
        dom = max(per_kernel, key=lambda k: per_kernel[k]["ms_per_launch"] * per_kernel[k]["launches"]) if per_kernel else None
        roofline = None
        if dom:
            rot_per_launch = rot_rank / per_kernel[dom]["launches"]
            tr = None
            if traffic and traffic.get("box") == box and dom in traffic.get("kernels", {}):
                # ncu capture at 32 rotations per launch (the internal chunk size); scaled to this run's average launch
                tr = traffic["kernels"][dom]["dram_bytes_per_launch"] * rot_per_launch / traffic.get("rotations_per_launch", 32)
            roofline = {"bound": "hbm", "kernel": dom, "achieved": per_kernel[dom]["algorithmic_GBps"], "peak": peak,
                        "unit": "GB/s", "frac": per_kernel[dom]["frac_of_hbm_peak"], "traffic": tr, "peak_source": peak_src,
                        "algorithmic_bytes_per_rotation": alg[dom], "rotations_per_launch": rot_per_launch,
                        "algorithmic_bytes_per_launch": alg[dom] * rot_per_launch,
                        "traffic_source": (traffic or {}).get("source") if tr else None,
                        "ncu_limiter": "FP32 issue + shared-memory pipe, not DRAM (profiles/): see roofline_fp32 for that ceiling"}
        step_gbs = (value / world) * alg["total"] / 1e9
        fk = "fp32" if args.filter_kind == "fp32" else "tcgen05"
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        lane_peak = 148 * 128 * sm_mhz * 1e6
        lane_ops = (value / world) / 2.0 * fp32_lane_ops_per_pair(box, fk)
        line = {
            "metric": "rotations/sec (all translations scored)", "value": value, "unit": "rotations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong" if full_set else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world, args.config),
            "step": {"rotations_per_step_per_gpu": nb, "rotations_total": n_done, "full_rotation_set": full_set,
                     "parallelism": ("rotations sharded i mod %d" % world) if not multi_pair else
                                    ("pairs sharded p mod %d, all rotations of a pair on one GPU" % world),
                     "timed_region": "K steps + all-gather of the per-rotation best poses + global best pose" if full_set else "K steps + all-gather"},
            "e2e": {"value": e2e_value, "unit": "rotations/s", "h2d_bytes_per_step": nb * (72 if atoms_path else 36),
                    "d2h_bytes_per_step": nb * 16, "call": "dlpd_search_host_atoms" if atoms_path else "dlpd_search_host",
                    "poses_identical_to_device_leg": e2e_identical},
            "gpu_launches": launches,
            "setup_ms": setup_ms,
            "roofline": roofline,
            "roofline_step": {"bound": "hbm", "achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak,
                              "algorithmic_bytes_per_rotation": alg["total"], "note": "whole step per GPU (SURVEY.md 8d)"},
            "roofline_fp32": {"achieved": lane_ops / 1e12, "peak": lane_peak / 1e12, "unit": "T lane-op/s", "frac": lane_ops / lane_peak,
                              "sm_mhz": sm_mhz, "lane_ops_per_rotation_pair": fp32_lane_ops_per_pair(box, fk),
                              "note": "FP32 lane-operations (model, DESIGN.md 3) / (148 SMs x 128 lanes x sampled SM clock); only "
                                      "packed FFMA2 / FADD2 reach that rate"},
            "kernels": per_kernel,
            "clocks": clocks,
        }
        if full_set:
            line["best_pose"] = {"rotation": int(result[1]), "score": float(gpu_all[int(result[1]), 0]),
                                 "xyz": [int(v) for v in gpu_all[int(result[1]), 1:]]}
        if world == 1 and args.cpu_rotations > 0 and atoms_path:
            poses = []
            t, n = cpu_baseline(rec_vols, rec_clash, lig_vols, lig_coords, filt, Rall, args.cpu_rotations, poses=poses)
            line["cpu_baseline"] = {"value": n / t, "unit": "rotations/s", "cores": os.cpu_count() or 1, "kind": "port",
                                    "sample": "%d rotations of oim10 (after 1 warm-up), CPU PyTorch torch.fft oracle, "
                                              "receptor spectra precomputed" % n}
            t2, n2 = cpu_baseline(rec_vols, rec_clash, lig_vols, lig_coords, filt, Rall, max(2, args.cpu_rotations // 4),
                                  retransform_receptor=True)
            line["cpu_baseline"]["value_receptor_retransformed_each_call"] = n2 / t2       # faithful to DockingModels.py:71
            if full_set:
                line["parity"] = parity_block(poses, gpu_all, rec_vols, rec_clash, lig_vols, lig_coords, filt, Rall)
                if line["parity"]["index_mismatches"] or line["parity"]["max_rel_score_err"] > line["parity"]["tol"]:
                    rc = 3
        if world == 1 and args.torch_gpu_rotations > 0 and atoms_path:
            eng.close()
            torch.cuda.empty_cache()
            try:
                line["gpu_torch_baseline"] = _torch_gpu_line(args, rec_vols, rec_clash, lig_vols, lig_coords.to(device), Rall, filt, device)
            except Exception as e:                       # a comparator must never cost the bench line
                line["gpu_torch_baseline"] = {"error": "%s: %s" % (type(e).__name__, e)}
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    if rc:
        sys.exit(rc)                                     # parity failure on the bench's own inputs: loud


def _torch_gpu_line(args, rec_vols, rec_clash, lig_vols, lig_dev, Rall, filt, device):
    """`gpu_torch_baseline` of the bench line: torch_loop_body timed on the same GPU, receptor re-transformed per call
    (faithful to VolumeConvolution) and with precomputed receptor spectra."""
    Rt = Rall[:args.torch_gpu_rotations + 1]
    cl = clash_volumes(lig_dev, Rt.to(device), device)
    n = args.torch_gpu_rotations
    return {
        "unit": "rotations/s", "kind": "port",
        "what": "the reference's loop body as plain PyTorch library calls on the same GPU (grid_sample, cuFFT "
                "rfftn/irfftn, interpolate, Linear, min): same-GPU comparator, SURVEY.md 8(d)",
        "receptor_retransformed_each_call": torch_gpu_baseline(rec_vols, rec_clash, lig_vols, cl, Rt, filt, n, False, device),
        "receptor_spectra_precomputed": torch_gpu_baseline(rec_vols, rec_clash, lig_vols, cl, Rt, filt, n, True, device),
        "sample": "%d rotations of oim10 after 1 warm-up, one rotation per call" % n}


if __name__ == "__main__":
    main()
