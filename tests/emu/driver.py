"""Host EMULATION of the CUDA kernel bodies -- TEST INFRASTRUCTURE ONLY.

`libdlpd_emu.so` is `csrc/dlpd_api.cu` compiled by g++ with -DDLPD_EMU: every "parallel for over tasks" of a
kernel body runs sequentially on the CPU, "device" pointers are host pointers.  It lets the `-m "not gpu"`
suite check index math, twiddles, layouts and the pair-packing algebra against the oracle without a GPU.
The package never loads it; GPU parity tests (`-m gpu`) go through the real libdlpd.so.
"""
import ctypes
import os
import subprocess

import torch

from deeplocalproteindocking_b200 import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(ROOT, "deeplocalproteindocking_b200", "csrc")
SO = os.path.join(HERE, "libdlpd_emu.so")


def build(force=False, flags=(), so=None):
    """flags / so: a second build of the emulation with other -D switches (e.g. the 8-CTA cluster code at L = 16)."""
    so = so or SO
    srcs = [os.path.join(SRC, f) for f in os.listdir(SRC)] + [os.path.join(ROOT, "include", "dlpd.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        # -DDLPD_EXPERIMENTAL=1: the emulation also covers the rejected epilogue variants the shipped library leaves out
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-DDLPD_EMU", "-DDLPD_EXPERIMENTAL=1", *flags, "-x", "c++", "-fPIC", "-shared",
                               "-o", so, os.path.join(SRC, "dlpd_api.cu")])
    return so


def use_variant(flags, tag):
    """Load a variant build as THE emulation library of this process; returns the previous one (pass it to `restore`)."""
    global _emu
    prev = _emu
    _emu = _lib.bind_signatures(ctypes.CDLL(build(flags=flags, so=os.path.join(HERE, "libdlpd_emu_%s.so" % tag))))
    return prev


def restore(prev):
    global _emu
    _emu = prev


_emu = None


def lib():
    global _emu
    if _emu is None:
        _emu = _lib.bind_signatures(ctypes.CDLL(build()))
    return _emu


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def fptr(t):
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_float))


class EmuPlan:
    def __init__(self, L, C0, C1, H, max_rot=4, center="grid"):
        self.L, self.C0, self.C1, self.H = L, C0, C1, H
        self.h = ctypes.c_void_p()
        _lib.check(lib(), lib().dlpd_plan_create(ctypes.byref(self.h), L, C0, C1, H, max_rot, _lib.CENTER_MODES[center], 0))

    def close(self):
        if self.h:
            lib().dlpd_plan_destroy(self.h)
            self.h = ctypes.c_void_p()

    def set_receptor(self, v0, v1, clash):
        v0, v1, clash = (t.contiguous().float() for t in (v0, v1, clash))
        _lib.check(lib(), lib().dlpd_set_receptor(self.h, ptr(v0), ptr(v1), ptr(clash), None))

    def set_ligand(self, v0, v1, clash=None):
        v0, v1 = v0.contiguous().float(), v1.contiguous().float()
        clash = clash.contiguous().float() if clash is not None else None
        _lib.check(lib(), lib().dlpd_set_ligand(self.h, ptr(v0), ptr(v1), ptr(clash), None))

    def set_filter(self, W1, b1, W2, b2):
        W1, b1, W2, b2 = (t.contiguous().float().reshape(-1) for t in (W1, b1, W2, b2))
        _lib.check(lib(), lib().dlpd_set_filter(self.h, fptr(W1), fptr(b1), fptr(W2), fptr(b2)))

    def search(self, R, tau, clash_rot=None, want_scores=False):
        R = R.contiguous().float().reshape(-1, 9)
        n, N = R.shape[0], 2 * self.L
        best = torch.zeros(n, 4)
        scores = torch.zeros(n, N, N, N) if want_scores else None
        clash_rot = clash_rot.contiguous().float() if clash_rot is not None else None
        _lib.check(lib(), lib().dlpd_search(self.h, ptr(R), ptr(clash_rot), n, float(tau), ptr(best), ptr(scores), None))
        return best, scores

    def search_candidates(self, R, tau, score_max, cap, clash_rot=None):
        R = R.contiguous().float().reshape(-1, 9)
        best = torch.zeros(R.shape[0], 4)
        cand = torch.zeros(cap, 4, dtype=torch.int32)
        count = torch.zeros(1, dtype=torch.int32)
        clash_rot = clash_rot.contiguous().float() if clash_rot is not None else None
        _lib.check(lib(), lib().dlpd_search_candidates(self.h, ptr(R), ptr(clash_rot), R.shape[0], float(tau), float(score_max),
                                                       ptr(best), ptr(cand), cap, ptr(count), None))
        total = int(count)
        c = cand[:min(total, cap)]
        return best, c[:, 0].contiguous().view(torch.float32), c[:, 1], c[:, 2], total

    def search_prerotated(self, v0, v1, clash, tau, want_scores=False):
        v0, v1, clash = (t.contiguous().float() for t in (v0, v1, clash))
        n, N = v0.shape[0], 2 * self.L
        best = torch.zeros(n, 4)
        scores = torch.zeros(n, N, N, N) if want_scores else None
        _lib.check(lib(), lib().dlpd_search_prerotated(self.h, ptr(v0), ptr(v1), ptr(clash), n, float(tau), ptr(best), ptr(scores), None))
        return best, scores

    def search_host(self, R, tau):
        R = R.contiguous().float().reshape(-1, 9)
        best = torch.zeros(R.shape[0], 4)
        _lib.check(lib(), lib().dlpd_search_host(self.h, ptr(R), R.shape[0], float(tau), ptr(best), None))
        return best

    def set_option(self, key, value):
        _lib.check(lib(), lib().dlpd_plan_set_option(self.h, key, value))

    def set_ligand_atoms(self, coords, res):
        coords = coords.contiguous().double().reshape(-1, 3)
        _lib.check(lib(), lib().dlpd_set_ligand_atoms(self.h, ptr(coords), coords.shape[0], float(res), None))

    def search_atoms(self, R64, tau, want_scores=False):
        R64 = R64.contiguous().double().reshape(-1, 9)
        n, N = R64.shape[0], 2 * self.L
        best = torch.zeros(n, 4)
        scores = torch.zeros(n, N, N, N) if want_scores else None
        _lib.check(lib(), lib().dlpd_search_atoms(self.h, ptr(R64), n, float(tau), ptr(best), ptr(scores), None))
        return best, scores

    def search_host_atoms(self, R64, tau):
        R64 = R64.contiguous().double().reshape(-1, 9)
        best = torch.zeros(R64.shape[0], 4)
        _lib.check(lib(), lib().dlpd_search_host_atoms(self.h, ptr(R64), R64.shape[0], float(tau), ptr(best), None))
        return best


OPT_PLANE_MODE, OPT_EPILOGUE_TILE, OPT_EPILOGUE_KIND = 1, 2, 3
EPILOGUE_AUTO, EPILOGUE_CLASSIC, EPILOGUE_PIPELINED, EPILOGUE_DUO, EPILOGUE_FUSED = 0, 1, 2, 3, 4
PLANE_AUTO, PLANE_SINGLE, PLANE_CLUSTER = 0, 1, 2


def set_option(key, value):
    _lib.check(lib(), lib().dlpd_set_option(key, value))


def volume_rotate(vol, R, center="grid"):
    vol, R = vol.contiguous().float(), R.contiguous().float()
    out = torch.empty_like(vol)
    B, C, L = vol.shape[:3]
    _lib.check(lib(), lib().dlpd_volume_rotate(ptr(vol), ptr(R), ptr(out), B, C, L, _lib.CENTER_MODES[center], None))
    return out


def volume_correlate(v1, v2):
    v1, v2 = v1.contiguous().float(), v2.contiguous().float()
    B, C, L = v1.shape[:3]
    out = torch.empty(B, C, 2 * L, 2 * L, 2 * L)
    _lib.check(lib(), lib().dlpd_volume_correlate(ptr(v1), ptr(v2), ptr(out), B, C, L, None))
    return out


def project_typed(coords, nat, off, L, res):
    coords = coords.contiguous().float()
    nat, off = nat.contiguous().to(torch.int32), off.contiguous().to(torch.int32)
    B, T = nat.shape
    out = torch.empty(B, T, L, L, L)
    _lib.check(lib(), lib().dlpd_project_typed(ptr(coords), ptr(nat), ptr(off), B, T, coords.shape[1] // 3, L, float(res), ptr(out), None))
    return out


def project_clash(coords, R, L, res):
    coords, R = coords.contiguous().double(), R.contiguous().double().reshape(-1, 9)
    out = torch.empty(R.shape[0], L, L, L)
    _lib.check(lib(), lib().dlpd_project_clash(ptr(coords), ptr(R), coords.shape[0], R.shape[0], L, float(res), ptr(out), None))
    return out
