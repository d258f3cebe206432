"""ctypes binding of the C ABI in `include/dlpd.h` (libdlpd.so, hand-written sm_100a CUDA).

There is no CPU fallback: `get_lib()` raises if the in-tree shared library is missing or cannot be
loaded, and every operator in this package goes through it.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_longlong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# DLPD_LIB_PATH selects another BUILD of the same CUDA library (tuning experiments); there is still no CPU path
LIB_PATH = os.environ.get("DLPD_LIB_PATH") or os.path.join(_HERE, "libdlpd.so")

# name -> (restype, argtypes); mirrors include/dlpd.h one to one
SIGNATURES = {
    "dlpd_version": (c_int, []),
    "dlpd_last_error": (c_char_p, []),
    "dlpd_search_supported": (c_int, [c_int, c_int, c_int, c_int]),
    "dlpd_op_supported": (c_int, [c_int]),
    "dlpd_plan_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "dlpd_plan_destroy": (c_int, [c_void_p]),
    "dlpd_plan_bytes": (c_longlong, [c_void_p]),
    "dlpd_set_receptor": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dlpd_set_ligand": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dlpd_set_filter": (c_int, [c_void_p, POINTER(c_float), POINTER(c_float), POINTER(c_float), POINTER(c_float)]),
    "dlpd_search": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "dlpd_search_candidates": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_void_p, c_void_p, c_int,
                                       c_void_p, c_void_p]),
    "dlpd_search_prerotated": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "dlpd_search_host": (c_int, [c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p]),
    "dlpd_set_ligand_atoms": (c_int, [c_void_p, c_void_p, c_int, c_float, c_void_p]),
    "dlpd_search_atoms": (c_int, [c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "dlpd_search_candidates_atoms": (c_int, [c_void_p, c_void_p, c_int, c_float, c_float, c_void_p, c_void_p, c_int, c_void_p,
                                             c_void_p]),
    "dlpd_search_host_atoms": (c_int, [c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p]),
    "dlpd_project_typed": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    "dlpd_plan_set_option": (c_int, [c_void_p, c_int, c_int]),
    "dlpd_build_flags": (c_int, []),
    "dlpd_release_workspace": (c_int, []),
    "dlpd_profile_enable": (c_int, [c_void_p, c_int]),
    "dlpd_profile_read": (c_int, [c_void_p, POINTER(ctypes.c_double), POINTER(c_int), c_void_p]),
    "dlpd_launch_count": (c_longlong, []),
    "dlpd_project_clash": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    "dlpd_set_option": (c_int, [c_int, c_int]),
    "dlpd_volume_rotate": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "dlpd_volume_correlate": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
}

CENTER_MODES = {"box": 0, "grid": 1}
OPTIONS = {"plane_mode": 1, "epilogue_tile": 2, "epilogue_kind": 3, "filter_kind": 4}      # DLPD_OPT_* of include/dlpd.h
EPILOGUE_KINDS = {"auto": 0, "classic": 1, "pipelined": 2, "duo": 3, "fused": 4, "tc_duo": 5, "pipe_tc": 6}
PLANE_MODES = {"auto": 0, "single": 1, "cluster": 2}
FILTER_KINDS = {"auto": 0, "fp32": 1, "tensor": 2, "tcgen05": 3}
BUILD_EXPERIMENTAL, BUILD_EMULATION = 1, 2


class DlpdError(RuntimeError):
    pass


def bind_signatures(cdll: ctypes.CDLL) -> ctypes.CDLL:
    """Attach restype/argtypes for every symbol of include/dlpd.h; raises AttributeError if one is missing."""
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(cdll, name)
        fn.restype = res
        fn.argtypes = args
    return cdll


def check(cdll: ctypes.CDLL, rc: int) -> None:
    """Error convention of the reference is a bare `raise Exception(...)` (Rotations.py:41); we raise DlpdError."""
    if rc != 0:
        msg = cdll.dlpd_last_error()
        raise DlpdError("libdlpd error %d: %s" % (rc, msg.decode() if msg else "?"))


_lib = None


def get_lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DlpdError(
                "libdlpd.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        _lib = bind_signatures(ctypes.CDLL(LIB_PATH))
    return _lib


def option_value(name: str, value) -> int:
    if name == "plane_mode" and isinstance(value, str):
        value = PLANE_MODES[value]
    if name == "epilogue_kind" and isinstance(value, str):
        value = EPILOGUE_KINDS[value]
    if name == "filter_kind" and isinstance(value, str):
        value = FILTER_KINDS[value]
    return int(value)


def set_option(name: str, value) -> None:
    """DEFAULT kernel-selection options that plans created afterwards start from (`dlpd_set_option`; a live plan is changed
    with `DockingSearch.set_option`): plane_mode = auto | single | cluster, epilogue_tile = 0 (auto) | 1 | 2,
    epilogue_kind = auto | classic (| pipelined | duo | fused in DLPD_EXPERIMENTAL builds), filter_kind = auto | fp32 | tcgen05
    (| tensor in experimental builds).  Tuning and tests only; the defaults are what ships."""
    lib = get_lib()
    check(lib, lib.dlpd_set_option(OPTIONS[name], option_value(name, value)))


def has_experimental() -> bool:
    """True if this libdlpd.so was built with -DDLPD_EXPERIMENTAL=1 (rejected epilogue variants, mma.sync filter)."""
    return bool(get_lib().dlpd_build_flags() & BUILD_EXPERIMENTAL)
