"""`Docker`: the search driver with the reference's call surface (`/root/reference/src/Docker/Docker.py`).

Same constructor arguments, `new_log` / `dockSE3` / `update_top`-ordering / `write_conformations` behaviour and
pose-file format; the rotation loop (Docker.py:211-236) is one `DockingSearch.search` call per chunk of
rotations instead of ~10 operator launches plus K host round trips per rotation.

Proteins are given as `Atoms` (coords [Nat,3] Angstrom float64, types [Nat] in 0..10) -- PDB parsing and atom
typing (`PDB2CoordsUnordered`, `Coords2TypedCoords`) are outside the search path; a `pdb_loader` callable can be
supplied to accept paths like the reference does.
"""
from __future__ import annotations

import os
import weakref
from dataclasses import dataclass
from typing import Callable, List, Optional, Tuple

import torch

from .rotations import Rotations
from .search import DockingSearch
from .volume import TypedCoords2Volume, project_clash_rotated

NUM_ATOM_TYPES = 11


@dataclass
class Atoms:
    coords: torch.Tensor      # [Nat,3] float64, Angstrom
    types: torch.Tensor       # [Nat] int64 in [0, NUM_ATOM_TYPES)

    def centred(self) -> "Atoms":
        """bbox centre to the origin (`load_batch(bbox_center=False)`, Docker.py:54-59)."""
        c = self.coords.to(torch.float64)
        return Atoms(c - 0.5 * (c.min(dim=0).values + c.max(dim=0).values), self.types)


def random_rotation(generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Uniform random rotation [3,3] float64 (stands in for TPL `getRandomRotation`, Docker.py:44)."""
    q = torch.randn(4, dtype=torch.float64, generator=generator)
    w, x, y, z = (q / q.norm()).tolist()
    return torch.tensor([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                         [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], dtype=torch.float64)


def typed_layout(atoms: Atoms, device) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Atoms -> (coords [1,3*Nat] sorted by type, num_atoms_of_type [1,11], offsets [1,11]): TPL's typed-coords layout."""
    order = torch.argsort(atoms.types, stable=True)
    coords = atoms.coords[order].reshape(1, -1)
    counts = torch.bincount(atoms.types, minlength=NUM_ATOM_TYPES)[:NUM_ATOM_TYPES]
    offsets = torch.cumsum(counts, 0) - counts
    return coords.to(device), counts.reshape(1, -1).to(torch.int32), offsets.reshape(1, -1).to(torch.int32)


def project_density_batched(coords: torch.Tensor, box_size: int, resolution: float) -> torch.Tensor:
    """[n,Nat,3] float32 -> [n,L,L,L]: summed (all types) Gaussian density, same stencil as TypedCoords2Volume.
    This is `project(...).sum(dim=1)` of the in-loop ligand re-projection (Docker.py:221-224), batched over rotations."""
    n, nat, _ = coords.shape
    L = box_size
    d = torch.arange(-2, 3, device=coords.device)
    stencil = torch.stack(torch.meshgrid(d, d, d, indexing="ij"), dim=-1).reshape(1, 1, -1, 3)
    ijk = torch.floor(coords / resolution).to(torch.int64).unsqueeze(2) + stencil                  # [n,Nat,125,3]
    ok = ((ijk >= 0) & (ijk < L)).all(dim=-1)
    r2 = ((coords.unsqueeze(2) - ijk.to(coords.dtype) * resolution) ** 2).sum(dim=-1)
    val = torch.exp(-r2 / 2.0) * ok.to(coords.dtype)
    c = ijk.clamp(0, L - 1)
    lin = (c[..., 0] * L + c[..., 1]) * L + c[..., 2]
    out = torch.zeros(n, L * L * L, device=coords.device, dtype=coords.dtype)
    out.scatter_add_(1, lin.reshape(n, -1), val.reshape(n, -1))
    return out.reshape(n, L, L, L)


def top_conformations(V: torch.Tensor, K: int) -> List[Tuple[int, int, int, float]]:
    """The K-iteration extraction loop of `Docker.update_top` (Docker.py:88-98) without K host round trips.

    Negative scores come out in (score, row-major index) order; once the minimum is >= 0 the reference's loop
    keeps returning the first zero voxel (masked, or already extracted and overwritten with 0.0), which this
    reproduces, duplicates included."""
    N1, N2, N3 = V.shape
    flat = V.reshape(-1)
    vals, idx = torch.sort(flat, stable=True)
    n_neg = int((vals < 0).sum())
    take = min(K, n_neg)
    v_h, i_h = vals[:take].tolist(), idx[:take].tolist()
    top = [(i // (N2 * N3), (i // N3) % N2, i % N3, s) for s, i in zip(v_h, i_h)]
    if len(top) == K:
        return top
    zeros = torch.nonzero(flat == 0).reshape(-1)
    pool = ([int(zeros[0])] if zeros.numel() else []) + i_h
    if not pool:
        i = int(idx[0])
        top.append((i // (N2 * N3), (i // N3) % N2, i % N3, float(vals[0])))
        pool = [i]
    i0 = min(pool)
    s0 = float(flat[i0]) if (zeros.numel() and i0 == int(zeros[0])) else 0.0
    while len(top) < K:
        top.append((i0 // (N2 * N3), (i0 // N3) % N2, i0 % N3, s0))
        s0 = 0.0
    return top


class Docker:
    def __init__(self, docking_model, angle_inc=15.0, box_size=80, resolution=1.25, max_conf=1000, randomize_rot=False,
                 rot_batch: int = 32, center: str = "grid", clash_source: str = "project",
                 pdb_loader: Optional[Callable[[str], Atoms]] = None, device=None, generator=None, verbose=False):
        """Reference arguments first (Docker.py:18); then: `rot_batch` rotations per fused call, rotation `center`
        convention ('grid' = what TorchProteinLibrary's grid_sample-based VolumeRotation does, default; 'box' opt-in),
        `clash_source` = 'project' (re-project rotated atoms, Docker.py:221-224) or 'rotate' (rotate the summed ligand
        density with the feature volumes)."""
        self.docking_model = docking_model
        self.log = None
        self.box_size, self.resolution = int(box_size), float(resolution)
        self.box_length = box_size * resolution
        self.max_conf = int(max_conf)
        self.rot = Rotations(angle_inc=angle_inc, verbose=verbose)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.project = TypedCoords2Volume(self.box_size, self.resolution)
        self.box_center = torch.full((1, 3), self.box_length / 2.0, dtype=torch.float64)
        self.randomize_rot = randomize_rot
        if randomize_rot:
            self.randR = random_rotation(generator).unsqueeze(0)
            if verbose:
                print("Adding random rotation to the receptor:", self.randR)
        self.rot_batch, self.center, self.clash_source = int(rot_batch), center, clash_source
        self.pdb_loader = pdb_loader
        self.top_list: list = []
        self.best_per_rotation: Optional[torch.Tensor] = None
        self._engine: Optional[DockingSearch] = None
        self._engine_key = None
        # The reference registers `cleanup` with atexit (Docker.py:47), which would keep every Docker -- and the multi-GB
        # device workspace of its search plan -- alive until interpreter exit.  A finalizer on the state that needs
        # closing does the same job without pinning the object.
        self._state = {"log": None, "engine": None}
        self._finalizer = weakref.finalize(self, Docker._release, self._state)

    @staticmethod
    def _release(state):
        if state.get("log") is not None:
            state["log"].close()
            state["log"] = None
        if state.get("engine") is not None:
            state["engine"].close()
            state["engine"] = None

    def close(self):
        """Close the log file and free the search plan's device memory (the object stays usable: a new plan is created on
        the next dock call)."""
        self.cleanup()
        if self._engine is not None:
            self._engine.close()
        self._engine = self._engine_key = None
        self._state["engine"] = None

    # ---- log handling: same resume rule as Docker.new_log (Docker.py:63-80) -----------------------------------
    def new_log(self, log_file_name, rewrite=True):
        if self.log is not None:
            self.log.close()
            self.log = None
        if os.path.exists(log_file_name) and not rewrite:
            with open(log_file_name, "r") as fin:
                n_lines = sum(1 for line in fin if len(line.split()) > 0)
            if n_lines > 1:
                return False
        self.log = open(log_file_name, "w")
        self._state["log"] = self.log
        return True

    def cleanup(self):
        if self.log is not None:
            self.log.close()
            self.log = None
        self._state["log"] = None

    # ---- set-up (Docker.py:190-209) ----------------------------------------------------------------------
    def _atoms(self, protein) -> Atoms:
        if isinstance(protein, Atoms):
            return protein
        if isinstance(protein, str):
            if self.pdb_loader is None:
                raise Exception("PDB input needs a pdb_loader (PDB parsing is outside the search path)", protein)
            return self.pdb_loader(protein)
        coords, types = protein
        return Atoms(torch.as_tensor(coords, dtype=torch.float64), torch.as_tensor(types, dtype=torch.int64))

    def _volumes(self, atoms: Atoms):
        coords, nat, off = typed_layout(Atoms(atoms.coords.to(torch.float32), atoms.types), self.device)
        density = self.project(coords, nat, off)                         # [1,11,L,L,L]
        vols = self._represent(density)
        return [v[0] for v in vols], density.sum(dim=1)[0]

    def _represent(self, density):
        """The representation network in true fp32 (the reference predates TF32 convolutions; cuDNN's default TF32
        path would perturb the volumes at the 1e-3 level)."""
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            return self.docking_model.representation(density)

    def _engine_for(self, vols) -> DockingSearch:
        key = (self.box_size, vols[0].shape[0], vols[1].shape[0])
        if self._engine is None or self._engine_key != key:
            if self._engine is not None:
                self._engine.close()
            hidden = self.docking_model.filter.fc[0].out_features
            self._engine = DockingSearch(self.box_size, (key[1], key[2]), hidden, max_rot=self.rot_batch,
                                         center=self.center, device=self.device)
            self._engine_key = key
            self._state["engine"] = self._engine
        return self._engine

    # ---- the search ----------------------------------------------------------------------------------------
    def dockSE3(self, ureceptor, uligand, batch_size=None, rotation_subset: Optional[torch.Tensor] = None, write: bool = True):
        """Dock two proteins (Docker.py:184-238).  `batch_size` is accepted for call compatibility (the reference
        scores that many rotations per step; here `rot_batch` rotations go into every fused call)."""
        self.top_list = []
        self.docking_model.eval()
        with torch.no_grad():
            rec = self._atoms(ureceptor).centred()
            lig = self._atoms(uligand).centred()
            rcoords = rec.coords
            if self.randomize_rot:
                rcoords = rcoords @ self.randR[0].t()
            rec_vols, rec_forbidden = self._volumes(Atoms(rcoords + self.box_center, rec.types))
            lig_vols, lig_density = self._volumes(Atoms(lig.coords + self.box_center, lig.types))
            eng = self._engine_for(rec_vols)
            eng.set_receptor(rec_vols, rec_forbidden)
            project = self.clash_source == "project"
            eng.set_ligand(lig_vols, None if project else lig_density)
            if project:          # per-rotation clash volumes re-projected inside the library (Docker.py:221-224)
                eng.set_ligand_atoms(lig.coords, self.resolution)
            eng.set_filter(self.docking_model.filter)
            tau = float(self.docking_model.threshold_clash)
            rot_idx = torch.arange(len(self.rot)) if rotation_subset is None else torch.as_tensor(rotation_subset)
            best_all = torch.empty(rot_idx.numel(), 4, device=self.device)
            want_topk = self.max_conf > 1
            N = 2 * self.box_size
            # how the top-K list was maintained (diagnostics): batches replayed exactly on a score volume, batches served by the
            # epilogue's candidate list at the current K-th score, batches whose threshold had to be bootstrapped by probing,
            # candidate-buffer overflows (each costs one exact replay) and threshold probes (one fused search each)
            self.exact_batches = self.candidate_batches = self.bootstrap_batches = self.overflow_batches = self.probe_searches = 0
            cap = max(4 * self.max_conf, 1 << 17)      # 2 MB of candidates: the first batches after the bootstrap run on a loose threshold
            pending = None       # candidate batch in flight: (PendingCandidates, sel, Rb, i0)

            def exact(Rb, sel):
                """Replay the reference's extraction loop on materialised score volumes (Docker.py:86-105)."""
                if project:
                    best, scores = eng.search_atoms(Rb, tau, return_scores=want_topk)
                else:
                    best, scores = eng.search(Rb.to(torch.float32), tau, None, return_scores=want_topk)
                if want_topk:
                    self.exact_batches += 1
                    for j, r in enumerate(sel.tolist()):
                        self.update_top(scores[j], r)
                return best

            def finish(pend):
                """Merge a candidate batch that was launched earlier (its list came back through pinned memory meanwhile)."""
                h, sel_p, Rb_p, i0_p = pend
                best, cand, total = h.result()
                if total <= cand.shape[0]:
                    self._merge_candidates(cand, sel_p, N)
                    self.candidate_batches += 1
                else:                                                    # candidate buffer overflowed: redo this batch exactly
                    self.overflow_batches += 1
                    best = exact(Rb_p, sel_p)
                best_all[i0_p:i0_p + sel_p.numel()] = best

            # float64 like r_batch_cpu (Docker.py:214); uploaded ONCE: a pageable host-to-device copy per batch would make the host
            # wait for the stream at every batch and drain the launch queue
            R_all = self.rot.R[rot_idx].to(self.device)
            for i0 in range(0, rot_idx.numel(), self.rot_batch):
                sel = rot_idx[i0:i0 + self.rot_batch]
                Rb = R_all[i0:i0 + self.rot_batch]
                # Top-K maintenance (Docker.py:86-105).  Once the list is full and its K-th score is negative, only
                # translations scoring <= that value can enter it: the fused epilogue lists exactly those and the score
                # volume is never written.  The batch is LAUNCHED with the K-th score known at that moment and merged one
                # iteration later, while the next batch runs (a stale, i.e. larger, threshold only lists a superset).
                kth = self.top_list[-1][4] if (want_topk and len(self.top_list) == self.max_conf) else None
                Rq = Rb if project else Rb.to(torch.float32)
                if want_topk and kth is not None and kth < 0.0:
                    h = eng.search_candidates_async(Rq, kth, tau, None, capacity=cap, use_atoms=project)
                    if pending is not None:
                        finish(pending)
                    pending = (h, sel, Rb, i0)
                    continue
                if pending is not None:
                    finish(pending)
                    pending = None
                best = None
                if want_topk:
                    # list not full yet (first batch): no threshold is known.  Bisect one on the device instead of writing and
                    # sorting rot_batch x (2L)^3 score volumes: a threshold theta < 0 with K <= #{V <= theta} <= capacity makes
                    # the candidate list a superset of the batch's K best translations (Docker.py:100-105 keeps K).
                    found = self._bootstrap_threshold(eng, Rq, tau, project, cap)
                    if found is not None:
                        best, cand = found
                        self._merge_candidates(cand.cpu(), sel, N)
                        self.bootstrap_batches += 1
                if best is None:
                    best = exact(Rb, sel)
                best_all[i0:i0 + sel.numel()] = best
            if pending is not None:
                finish(pending)
            self.best_per_rotation = best_all.cpu()
            if not want_topk:
                b = self.best_per_rotation
                items = [(int(r), int(b[j, 1]), int(b[j, 2]), int(b[j, 3]), float(b[j, 0])) for j, r in enumerate(rot_idx.tolist())]
                items.sort(key=lambda t: t[4])
                self.top_list = items[:self.max_conf]
        if write:
            self.write_conformations()

    def _bootstrap_threshold(self, eng, Rq, tau, project, cap, max_probes: int = 14):
        """Find theta in (min score of the batch, 0) whose candidate list has between K - len(list) ... cap entries; every probe
        is one fused search (no score volume).  None if the batch has fewer negative scores than needed (the caller then
        replays the reference's extraction loop on a score volume, zeros and duplicates included)."""
        need = self.max_conf
        best = (eng.search_atoms(Rq, tau) if project else eng.search(Rq, tau))[0]
        self.probe_searches += 1
        lo = float(best[:, 0].min())                     # count(lo) >= 1
        if not lo < 0.0:
            return None
        hi, theta = 0.0, 0.5 * lo
        for _ in range(max_probes):
            best, cand, total = eng.search_candidates(Rq, theta, tau, None, capacity=cap, use_atoms=project)
            self.probe_searches += 1
            if need <= total <= cand.shape[0]:
                return best, cand
            if total < need:
                lo, theta = theta, 0.5 * (theta + hi)    # too few: move towards 0
                if theta > -1e-30:
                    return None                          # fewer than K negative scores in the whole batch
            else:
                hi, theta = theta, 0.5 * (theta + lo)    # too many for the buffer: move towards the minimum
        return None

    def _merge_candidates(self, cand: torch.Tensor, sel: torch.Tensor, N: int):
        """Merge (score, position in `sel`, linear index) candidates into the global list: same order as appending each
        rotation's extraction list and stable-sorting by score (Docker.py:100-105) -- ties keep rotation order, then
        row-major voxel order (Docker.py:91-96)."""
        if cand.shape[0] > self.max_conf:        # long list (loose threshold): only its K best rows can enter -- cut it with tensor sorts
            for col in (2, 1, 0):                # stable sorts, least significant key first = order by (score, rotation, voxel)
                cand = cand[torch.argsort(cand[:, col], stable=True)]
            cand = cand[:self.max_conf]
        rows = sorted(((float(s), int(j), int(lin)) for s, j, lin in cand.tolist()), key=lambda t: (t[0], t[1], t[2]))
        sel_l = sel.tolist()
        for s, j, lin in rows:
            self.top_list.append((sel_l[j], lin // (N * N), (lin // N) % N, lin % N, s))
        self.top_list.sort(key=lambda t: t[4])
        self.top_list = self.top_list[:self.max_conf]

    def dockE3(self, ureceptor, uligand, batch_size=None, rotation_subset: Optional[torch.Tensor] = None):
        """`Docker.dockE3` (Docker.py:135-182): the ligand ATOMS are rotated, re-projected and pushed through the
        representation network for every rotation (plain PyTorch; not rotation-equivariant networks need this), then
        the same fused correlation / filter / mask / argmin path scores all translations."""
        self.top_list = []
        self.docking_model.eval()
        with torch.no_grad():
            rec = self._atoms(ureceptor).centred()
            lig = self._atoms(uligand).centred()
            rcoords = rec.coords
            if self.randomize_rot:
                rcoords = rcoords @ self.randR[0].t()
            rec_vols, rec_forbidden = self._volumes(Atoms(rcoords + self.box_center, rec.types))
            eng = self._engine_for(rec_vols)
            eng.set_receptor(rec_vols, rec_forbidden)
            eng.set_filter(self.docking_model.filter)
            tau = float(self.docking_model.threshold_clash)
            rot_idx = torch.arange(len(self.rot)) if rotation_subset is None else torch.as_tensor(rotation_subset)
            order = torch.argsort(lig.types, stable=True)
            counts = torch.bincount(lig.types, minlength=NUM_ATOM_TYPES)[:NUM_ATOM_TYPES]
            lig_sorted = lig.coords[order]
            best_all = torch.empty(rot_idx.numel(), 4, device=self.device)
            nb = self.rot_batch if batch_size is None else max(int(batch_size), 1)
            for i0 in range(0, rot_idx.numel(), nb):
                sel = rot_idx[i0:i0 + nb]
                Rb = self.rot.R[sel]                                                    # CPU float64 (Docker.py:159,162)
                moved = torch.einsum("nij,aj->nai", Rb, lig_sorted) + self.box_center   # rotate, then translate (:162-163)
                n = sel.numel()
                coords = moved.reshape(n, -1).to(self.device, torch.float32)
                nat = counts.reshape(1, -1).repeat(n, 1).to(self.device, torch.int32)
                off = (torch.cumsum(counts, 0) - counts).reshape(1, -1).repeat(n, 1).to(self.device, torch.int32)
                ligand = self.project(coords, nat, off)                                 # [n,11,L,L,L]          (:164)
                vols = self._represent(ligand)                                          #                       (:165)
                best, scores = eng.search_prerotated(vols[0], vols[1], ligand.sum(dim=1), tau, return_scores=self.max_conf > 1)
                best_all[i0:i0 + n] = best
                if self.max_conf > 1:
                    for j, r in enumerate(sel.tolist()):
                        self.update_top(scores[j], r)
            self.best_per_rotation = best_all.cpu()
            if self.max_conf <= 1:
                b = self.best_per_rotation
                items = [(int(r), int(b[j, 1]), int(b[j, 2]), int(b[j, 3]), float(b[j, 0])) for j, r in enumerate(rot_idx.tolist())]
                items.sort(key=lambda t: t[4])
                self.top_list = items[:self.max_conf]
        self.write_conformations()

    def update_top(self, V, rotation_index):
        """Merge the `max_conf` best translations of one rotation into the global list (Docker.py:86-105)."""
        for x, y, z, score in top_conformations(V, self.max_conf):
            self.top_list.append((rotation_index, x, y, z, score))
        self.top_list.sort(key=lambda t: t[4])
        self.top_list = self.top_list[:self.max_conf]

    # ---- output (Docker.py:107-133) --------------------------------------------------------------------------
    def conformation(self, entry):
        i, x, y, z, score = entry
        r = self.rot.R[i]
        t = torch.tensor([a if a < self.box_size else -(2 * self.box_size - a) for a in (x, y, z)], dtype=torch.float64)
        t = t * self.resolution
        if self.randomize_rot:
            rt = self.randR[0].t()
            t, r = rt @ t, rt @ r
        return r, t, score

    def write_conformations(self):
        if self.log is None:
            return
        for entry in self.top_list:
            r, t, score = self.conformation(entry)
            vals = [float(r[a, b]) for a in range(3) for b in range(3)] + [float(v) for v in t]
            self.log.write("".join("%f\t" % v for v in vals) + "%f\n" % score)
        self.log.flush()
