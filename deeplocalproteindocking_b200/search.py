"""Fused exhaustive search engine: the rotation loop of `Docker.dockSE3`
(`/root/reference/src/Docker/Docker.py:211-236`) as one call per batch of rotations.

Per rotation the reference runs VolumeRotation x2, CPU re-projection of the ligand atoms,
VolumeConvolution x3, F.interpolate, cat/transpose, the SimpleFilter MLP, the clash mask and
`update_top`.  Here `DockingSearch.search` hands a batch of rotation matrices to libdlpd.so
(`include/dlpd.h: dlpd_search`), which keeps the receptor spectra resident, never materialises the
rotated, correlation or score volumes and returns the best translation of every rotation.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import torch

from . import _lib


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _fptr(t: torch.Tensor):
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_float))


def _stream_ptr(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _dev_f32(t: torch.Tensor, device, what: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % what)
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


class PendingCandidates:
    """Result of `DockingSearch.search_candidates_async`: the candidate list and its count travel to pinned host memory
    behind an event, so the caller can launch the next batch before it looks at this one."""

    def __init__(self, best, cand, count, capacity, host_cand, host_count, device):
        self.best, self.capacity = best, int(capacity)
        self._cand_h, self._count_h = host_cand, host_count
        self._cand_h.copy_(cand, non_blocking=True)
        self._count_h.copy_(count, non_blocking=True)
        self._ev = torch.cuda.Event()
        self._ev.record(torch.cuda.current_stream(device))

    def result(self) -> Tuple[torch.Tensor, torch.Tensor, int]:
        """(best [n,4] on the device, candidates [m,3] float64 on the HOST = (score, rotation index, linear index), total)."""
        self._ev.synchronize()
        total = int(self._count_h[0])
        c = self._cand_h[:min(total, self.capacity)]
        out = torch.stack([c[:, 0].contiguous().view(torch.float32).double(), c[:, 1].double(), c[:, 2].double()], dim=1)
        return self.best, out, total


class DockingSearch:
    """One receptor-ligand pair on one GPU.

    box_size L, channel counts (C0 @ L, C1 @ L/2) and hidden width H select a compiled kernel
    configuration (`dlpd_search_supported`); anything else raises -- there is no fallback path.
    `max_rot` rotations are processed per internal batch (two rotations share one complex transform).
    `center` is the rotation centre of the ligand feature volumes: 'grid' = (L-1)/2, what `affine_grid` + `grid_sample`
    (the construction TorchProteinLibrary's VolumeRotation is built on) does -- the default wherever drop-in parity is
    claimed; 'box' = L/2, consistent with voxel i at i * resolution and atoms centred at box_length / 2 (opt-in).
    """

    def __init__(self, box_size: int, channels: Sequence[int] = (16, 32), hidden: Optional[int] = None,
                 max_rot: int = 16, center: str = "grid", device=None):
        if not torch.cuda.is_available():
            raise _lib.DlpdError("DockingSearch needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.get_lib()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.L = int(box_size)
        self.C0, self.C1 = int(channels[0]), int(channels[1])
        self.H = int(hidden) if hidden is not None else (self.C0 + self.C1) // 2
        self.max_rot = int(max_rot)
        self.center = center
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib, self.lib.dlpd_plan_create(ctypes.byref(self._h), self.L, self.C0, self.C1, self.H,
                                                         self.max_rot, _lib.CENTER_MODES[center], self.device.index or 0))
        self.threshold_clash = 300.0
        self._keep = {}

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self.lib.dlpd_plan_destroy(self._h)     # restores the caller's current device itself
            self._h = ctypes.c_void_p()

    def set_option(self, name: str, value) -> None:
        """Kernel selection of THIS plan (`dlpd_plan_set_option`): see `_lib.set_option` for the names.  Changing the plane
        mode after `set_receptor` makes the next search fail with DLPD_ESTATE until the receptor is set again."""
        _lib.check(self.lib, self.lib.dlpd_plan_set_option(self._h, _lib.OPTIONS[name], _lib.option_value(name, value)))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.dlpd_plan_bytes(self._h))

    # -- per-pair state ----------------------------------------------------------------------------
    def _check_levels(self, vols, what):
        L, L1 = self.L, self.L // 2
        v0, v1 = vols
        v0 = v0.reshape(-1, L, L, L) if v0.dim() == 5 else v0
        v1 = v1.reshape(-1, L1, L1, L1) if v1.dim() == 5 else v1
        if tuple(v0.shape) != (self.C0, L, L, L) or tuple(v1.shape) != (self.C1, L1, L1, L1):
            raise ValueError("%s volumes must be [%d,%d^3] and [%d,%d^3], got %s and %s"
                             % (what, self.C0, L, self.C1, L1, tuple(v0.shape), tuple(v1.shape)))
        return _dev_f32(v0, self.device, what), _dev_f32(v1, self.device, what)

    def set_receptor(self, receptor_volumes, receptor_forbidden):
        """receptor_volumes = representation(receptor) (Docker.py:205), receptor_forbidden = receptor.sum(1) (:206)."""
        v0, v1 = self._check_levels(receptor_volumes, "receptor")
        f = _dev_f32(receptor_forbidden, self.device, "receptor_forbidden").reshape(self.L, self.L, self.L)
        with torch.cuda.device(self.device):
            _lib.check(self.lib, self.lib.dlpd_set_receptor(self._h, _ptr(v0), _ptr(v1), _ptr(f), _stream_ptr(self.device)))
        self._keep["rec"] = (v0, v1, f)

    def set_ligand(self, ligand_volumes, ligand_forbidden=None):
        """Un-rotated ligand volumes (Docker.py:208-209).  `ligand_forbidden` [L^3] is rotated with the same operator
        as the features; leave it None to pass re-projected per-rotation clash volumes to `search` (Docker.py:221-224)."""
        v0, v1 = self._check_levels(ligand_volumes, "ligand")
        f = None
        if ligand_forbidden is not None:
            f = _dev_f32(ligand_forbidden, self.device, "ligand_forbidden").reshape(self.L, self.L, self.L)
        with torch.cuda.device(self.device):
            _lib.check(self.lib, self.lib.dlpd_set_ligand(self._h, _ptr(v0), _ptr(v1), _ptr(f), _stream_ptr(self.device)))
        self._keep["lig"] = (v0, v1, f)

    def set_ligand_atoms(self, coords: torch.Tensor, resolution: float):
        """Bbox-centred ligand atoms [Nat,3] (float64, Angstrom) for the in-loop clash re-projection inside the library
        (Docker.py:221-224): `search_atoms` / `search_host_atoms` then need nothing but the float64 rotation matrices."""
        c = coords.detach().to(device=self.device, dtype=torch.float64).contiguous().reshape(-1, 3)
        with torch.cuda.device(self.device):
            _lib.check(self.lib, self.lib.dlpd_set_ligand_atoms(self._h, _ptr(c), c.shape[0], float(resolution),
                                                               _stream_ptr(self.device)))
        self._keep["atoms"] = c

    def set_filter(self, filt):
        """`filt`: a SimpleFilter-like module (`.fc[0]`, `.fc[2]` Linear layers, DockingModels.py:28-32) or (W1,b1,W2,b2)."""
        if isinstance(filt, (tuple, list)):
            W1, b1, W2, b2 = filt
        else:
            W1, b1, W2, b2 = filt.fc[0].weight, filt.fc[0].bias, filt.fc[2].weight, filt.fc[2].bias
        host = [t.detach().to("cpu", torch.float32).contiguous().reshape(-1) for t in (W1, b1, W2, b2)]
        if host[0].numel() != self.H * (self.C0 + self.C1) or host[1].numel() != self.H or host[2].numel() != self.H:
            raise ValueError("filter shape does not match the plan (H=%d, inputs=%d)" % (self.H, self.C0 + self.C1))
        _lib.check(self.lib, self.lib.dlpd_set_filter(self._h, *[_fptr(t) for t in host]))
        self._keep["filt"] = host

    # -- the hot call --------------------------------------------------------------------------------
    def search(self, R: torch.Tensor, threshold_clash: Optional[float] = None,
               ligand_forbidden_rotated: Optional[torch.Tensor] = None, return_scores: bool = False
               ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """R [n,3,3] -> best [n,4] = (score, x, y, z) of the lowest masked score per rotation.

        With `return_scores` the masked score volumes V [n,2L,2L,2L] (Docker.py:232) are also written --
        the module-level drop-in path; the production search never materialises them."""
        tau = self.threshold_clash if threshold_clash is None else float(threshold_clash)
        Rd = _dev_f32(R, self.device, "R").reshape(-1, 9)
        n, N = Rd.shape[0], 2 * self.L
        best = torch.empty(n, 4, device=self.device, dtype=torch.float32)
        scores = torch.empty(n, N, N, N, device=self.device, dtype=torch.float32) if return_scores else None
        cr = None
        if ligand_forbidden_rotated is not None:
            cr = _dev_f32(ligand_forbidden_rotated, self.device, "ligand_forbidden_rotated").reshape(n, self.L, self.L, self.L)
        with torch.cuda.device(self.device):
            _lib.check(self.lib, self.lib.dlpd_search(self._h, _ptr(Rd), _ptr(cr), n, tau, _ptr(best), _ptr(scores),
                                                     _stream_ptr(self.device)))
        return best, scores

    def search_atoms(self, R64: torch.Tensor, threshold_clash: Optional[float] = None, return_scores: bool = False
                     ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """`search` with the reference's own per-rotation clash volumes: R64 [n,3,3] float64 (the table's matrices,
        Docker.py:214); the library casts them to fp32 for the volume rotation (Docker.py:215) and re-projects the rotated
        atoms given to `set_ligand_atoms` chunk by chunk (Docker.py:221-224)."""
        tau = self.threshold_clash if threshold_clash is None else float(threshold_clash)
        Rd = R64.detach().to(device=self.device, dtype=torch.float64).contiguous().reshape(-1, 9)
        n, N = Rd.shape[0], 2 * self.L
        best = torch.empty(n, 4, device=self.device, dtype=torch.float32)
        scores = torch.empty(n, N, N, N, device=self.device, dtype=torch.float32) if return_scores else None
        with torch.cuda.device(self.device):
            _lib.check(self.lib, self.lib.dlpd_search_atoms(self._h, _ptr(Rd), n, tau, _ptr(best), _ptr(scores),
                                                           _stream_ptr(self.device)))
        return best, scores

    def search_host_atoms(self, R64_host: torch.Tensor, threshold_clash: Optional[float] = None) -> torch.Tensor:
        """End to end with HOST buffers: R64 [n,3,3] CPU float64 in, best [n,4] CPU float32 out; H2D, clash re-projection,
        fused search, D2H and a stream synchronisation inside the library (`dlpd_search_host_atoms`)."""
        tau = self.threshold_clash if threshold_clash is None else float(threshold_clash)
        Rh = R64_host.detach().to("cpu", torch.float64).contiguous().reshape(-1, 9)
        out = torch.empty(Rh.shape[0], 4, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib, self.lib.dlpd_search_host_atoms(self._h, _ptr(Rh), Rh.shape[0], tau, _ptr(out),
                                                                _stream_ptr(self.device)))
        return out

    def search_candidates(self, R: torch.Tensor, score_max: float, threshold_clash: Optional[float] = None,
                          ligand_forbidden_rotated: Optional[torch.Tensor] = None, capacity: int = 1 << 16,
                          use_atoms: bool = False) -> Tuple[torch.Tensor, torch.Tensor, int]:
        """`search` that also lists every translation with masked score <= `score_max` (the caller's current K-th best,
        Docker.py:100-105).  Returns (best [n,4], candidates [m,3] float64 = (score, rotation index in R, linear index),
        total) -- if total > m == capacity the list is truncated and the caller must redo this batch exactly."""
        best, out, total = self.search_candidates_async(R, score_max, threshold_clash, ligand_forbidden_rotated, capacity,
                                                        use_atoms).result()
        return best, out.to(self.device), total

    def search_candidates_async(self, R: torch.Tensor, score_max: float, threshold_clash: Optional[float] = None,
                                ligand_forbidden_rotated: Optional[torch.Tensor] = None, capacity: int = 1 << 16,
                                use_atoms: bool = False) -> PendingCandidates:
        """The same without waiting: returns a `PendingCandidates` whose `.result()` blocks until THIS batch is done.  Two
        pinned host buffers are used alternately, so at most two results may be outstanding."""
        tau = self.threshold_clash if threshold_clash is None else float(threshold_clash)
        if use_atoms:        # R float64: clash volumes re-projected inside the library from the atoms of set_ligand_atoms
            Rd = R.detach().to(device=self.device, dtype=torch.float64).contiguous().reshape(-1, 9)
        else:
            Rd = _dev_f32(R, self.device, "R").reshape(-1, 9)
        n = Rd.shape[0]
        best = torch.empty(n, 4, device=self.device, dtype=torch.float32)
        cand = torch.empty(int(capacity), 4, device=self.device, dtype=torch.int32)
        count = torch.zeros(1, device=self.device, dtype=torch.int32)
        cr = None
        if ligand_forbidden_rotated is not None:
            cr = _dev_f32(ligand_forbidden_rotated, self.device, "ligand_forbidden_rotated").reshape(n, self.L, self.L, self.L)
        with torch.cuda.device(self.device):
            if use_atoms:
                _lib.check(self.lib, self.lib.dlpd_search_candidates_atoms(self._h, _ptr(Rd), n, tau, float(score_max), _ptr(best),
                                                                          _ptr(cand), int(capacity), _ptr(count),
                                                                          _stream_ptr(self.device)))
            else:
                _lib.check(self.lib, self.lib.dlpd_search_candidates(self._h, _ptr(Rd), _ptr(cr), n, tau, float(score_max),
                                                                    _ptr(best), _ptr(cand), int(capacity), _ptr(count),
                                                                    _stream_ptr(self.device)))
        slots = self._keep.setdefault("cand_host", {})
        k = self._keep["cand_flip"] = 1 - self._keep.get("cand_flip", 0)
        if k not in slots or slots[k][0].shape[0] != int(capacity):
            slots[k] = (torch.empty(int(capacity), 4, dtype=torch.int32).pin_memory(), torch.empty(1, dtype=torch.int32).pin_memory())
        return PendingCandidates(best, cand, count, capacity, slots[k][0], slots[k][1], self.device)

    def search_prerotated(self, lig0_rot: torch.Tensor, lig1_rot: torch.Tensor, ligand_forbidden_rotated: torch.Tensor,
                          threshold_clash: Optional[float] = None, return_scores: bool = False
                          ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """The loop body of `Docker.dockE3` (Docker.py:159-180): ligand volumes [n,C0,L^3], [n,C1,(L/2)^3] and clash
        volumes [n,L^3] computed per rotation by the caller; returns what `search` returns."""
        tau = self.threshold_clash if threshold_clash is None else float(threshold_clash)
        L, L1 = self.L, self.L // 2
        n, N = lig0_rot.shape[0], 2 * self.L
        v0 = _dev_f32(lig0_rot, self.device, "lig0_rot").reshape(n, self.C0, L, L, L)
        v1 = _dev_f32(lig1_rot, self.device, "lig1_rot").reshape(n, self.C1, L1, L1, L1)
        cr = _dev_f32(ligand_forbidden_rotated, self.device, "ligand_forbidden_rotated").reshape(n, L, L, L)
        best = torch.empty(n, 4, device=self.device, dtype=torch.float32)
        scores = torch.empty(n, N, N, N, device=self.device, dtype=torch.float32) if return_scores else None
        with torch.cuda.device(self.device):
            _lib.check(self.lib, self.lib.dlpd_search_prerotated(self._h, _ptr(v0), _ptr(v1), _ptr(cr), n, tau, _ptr(best),
                                                                _ptr(scores), _stream_ptr(self.device)))
        return best, scores

    def search_host(self, R_host: torch.Tensor, threshold_clash: Optional[float] = None) -> torch.Tensor:
        """End-to-end call with HOST buffers: R [n,3,3] CPU float32 in, best [n,4] CPU out (copies inside, synchronous)."""
        tau = self.threshold_clash if threshold_clash is None else float(threshold_clash)
        Rh = R_host.detach().to("cpu", torch.float32).contiguous().reshape(-1, 9)
        out = torch.empty(Rh.shape[0], 4, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib, self.lib.dlpd_search_host(self._h, _ptr(Rh), Rh.shape[0], tau, _ptr(out),
                                                          _stream_ptr(self.device)))
        return out

    # -- measurement hooks (bench.py) ------------------------------------------------------------------
    PROFILE_KINDS = ("fwdz_level0", "fwdz_level1", "plane_level0", "plane_level1", "epilogue", "hidden_level1")

    def profile(self, on: bool = True):
        _lib.check(self.lib, self.lib.dlpd_profile_enable(self._h, 1 if on else 0))

    def profile_read(self):
        """{kernel kind: (total ms, launches)} measured with CUDA events on the launching stream; resets the counters."""
        ms = (ctypes.c_double * len(self.PROFILE_KINDS))()
        n = (ctypes.c_int * len(self.PROFILE_KINDS))()
        with torch.cuda.device(self.device):
            _lib.check(self.lib, self.lib.dlpd_profile_read(self._h, ms, n, _stream_ptr(self.device)))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(self.PROFILE_KINDS)}

    def launch_count(self) -> int:
        return int(self.lib.dlpd_launch_count())
