"""ctypes binding of libsiss_b200.so (the C ABI declared in include/siss_b200.h).

There is deliberately NO fallback: if the shared object is missing or the device is not a B200
(sm_100), every entry point raises. The CPU restatement of the algorithm lives under ``oracle/``
and is test infrastructure only.
"""
from __future__ import annotations

import ctypes
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_void_p, POINTER
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "libsiss_b200.so"

SISS_F32, SISS_BF16, SISS_F16 = 0, 1, 2
SISS_COMBINE_SCALING_NORM, SISS_COMBINE_ERASEDIFF, SISS_COMBINE_NONE = 0, 1, 2
ABI_VERSION = 3


class SissLibraryError(RuntimeError):
    """libsiss_b200.so is missing, stale, or returned an error code."""


# name -> (restype, argtypes); mirrors include/siss_b200.h one to one
_P, _I, _L, _F, _D = c_void_p, c_int, c_int64, c_float, c_double
_U = ctypes.c_uint64
SIGNATURES = {
    "siss_abi_version": (_I, []),
    "siss_error_string": (c_char_p, [_I]),
    "siss_check_device": (_I, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "siss_add_noise": (_I, [_P, _P, _P, _P, _I, _P, _L, _L, _I, _P]),
    "siss_add_noise_pair": (_I, [_P, _P, _P, _P, _P, _I, _P, _P, _L, _L, _I, _P]),
    "siss_row_workspace_bytes": (_L, [_L]),
    "siss_mixture_weights": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _D, _P, _P, _P, _P, _P, _P, _L, _L, _I, _P]),
    "siss_add_noise_mixture": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _D, _P, _P, _P, _P, _P, _P, _L, _L, _I, _P]),
    "siss_wmse_fwd_bwd": (_I, [_P, _I, _P, _P, _P, _I, _P, _P, _P, _I, _P, _P, _F, _F, _P, _P, _P, _P, _P, _L, _L, _P]),
    "siss_wmse_fwd": (_I, [_P, _I, _P, _P, _P, _I, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _L, _L, _P]),
    "siss_wmse_bwd": (_I, [_P, _I, _P, _P, _P, _I, _P, _P, _P, _I, _P, _P,
                           _P, _I, _P, _I, _P, _I, _P, _I, _P, _L, _L, _P]),
    "siss_sqerr_fwd": (_I, [_P, _I, _P, _I, _P, _P, _F, _L, _P]),
    "siss_sqerr_bwd": (_I, [_P, _I, _P, _I, _P, _I, _P, _I, _F, _I, _P, _L, _P]),
    "siss_dual_mse_fwd_bwd": (_I, [_P, _P, _I, _P, _P, _I, _F, _F, _P, _P, _P, _P, _P, _L, _L, _P]),
    "siss_norm3_workspace_bytes": (_L, []),
    "siss_norm3": (_I, [_P, _P, _L, _P, _P, _P]),
    "siss_combine": (_I, [_P, _P, _P, _L, _P, _I, _F, _F, _I, _P, _P]),
    "siss_mt_chunk_elems": (_I, []),
    "siss_mt_norm3": (_I, [_P, _P, _P, _P, _I, _L, _P, _P, _P]),
    "siss_mt_combine": (_I, [_P, _P, _P, _P, _P, _I, _L, _P, _I, _F, _F, _I, _P, _P]),
    "siss_combine_adamw": (_I, [_P, _P, _L, _P, _I, _F, _F, _I, _P, _P, _P, _D, _D, _D, _D, _D, _L, _P, _P, _P, _D, _I, _P, _P, _P]),
    "siss_counter_add": (_I, [_P, _L, _P]),
    "siss_randn": (_I, [_P, _L, _I, _U, _U, _P, _U, _P]),
    "siss_draw_rows": (_I, [_P, _P, _L, _U, _U, _P, _U, _L, _L, _D, _P]),
    "siss_add_noise_mixture_rng": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _D, _U, _U, _P, _U, _P, _P, _P, _P, _P, _P, _P, _L, _L, _I, _P]),
    "siss_dual_mse_rng_fwd_bwd": (_I, [_P, _P, _I, _P, _I, _U, _U, _P, _U, _F, _F, _P, _P, _P, _P, _P, _P, _L, _L, _P]),
    "siss_membership_add_noise": (_I, [_P, _P, _P, _P, _I, _L, _P, _P, _L, _L, _L, _L, _I, _P]),
    "siss_membership_sqerr": (_I, [_P, _P, _P, _I, _P, _P, _P, _L, _L, _L, _L, _P]),
    "siss_batch_stats": (_I, [_P, _P, _P, _P, _L, _L, _P, _P]),
    "siss_p2p_workspace_bytes": (_L, []),
    "siss_p2p_reduce_norm3": (_I, [_P, _P, _P, _I, _I, _L, _P, _P, _P, _I, _P, _P]),
    "siss_publish_sums": (_I, [_P, _I, _P, _I, _I, _P]),
    "siss_p2p_combine_allgather": (_I, [_P, _P, _P, _P, _I, _I, _L, _I, _F, _F, _I, _P, _P]),
    "siss_p2p_adamw_allgather": (_I, [_P, _P, _P, _P, _I, _I, _L, _I, _F, _F, _I, _P, _P, _D, _D, _D, _D, _D, _L, _P, _P,
                                      _P, _D, _P, _P]),
    "siss_nvls_reduce_norm3": (_I, [_P, _P, _P, _I, _I, _L, _P, _P, _P, _I, _P, _P]),
    "siss_nvls_combine_allgather": (_I, [_P, _P, _P, _P, _I, _I, _L, _I, _F, _F, _I, _P, _P]),
    "siss_nvls_adamw_allgather": (_I, [_P, _P, _P, _P, _P, _I, _I, _L, _I, _F, _F, _I, _P, _P, _D, _D, _D, _D, _D, _L, _P, _P,
                                       _P, _D, _P, _P]),
    "siss_nvls_xcombine_bcast": (_I, [_P, _P, _P, _P, _I, _I, _L, _F, _I, _P, _P]),
    "siss_scale_finalize": (_I, [_P, _L, _P, _P, _I, _F, _F, _I, _P, _P]),
    "siss_ce_reduce_norm3": (_I, [_P, _P, _P, _I, _I, _L, _P, _P, _P, _P, _I, _I, _P, _P]),
    "siss_ce_combine_allgather": (_I, [_P, _P, _P, _P, _I, _I, _L, _I, _I, _F, _F, _I, _P, _P]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the library once; raise SissLibraryError if it is absent or from another ABI."""
    global _lib
    if _lib is not None:
        return _lib
    # Not a fallback: the only thing tried is compiling the SAME CUDA library in place (needs nvcc). build() returns
    # at once when the library matches the content hash of its sources, so a kernel edit can never run a stale .so.
    from . import build as _build
    if _build.needs_build():
        try:
            _build.build()
        except Exception as e:
            what = "is stale (its sources changed since it was built)" if LIB_PATH.exists() else "not found"
            raise SissLibraryError(
                f"{LIB_PATH} {what} and could not be built ({e}). Build it with `python -m siss_b200.build` "
                "(or __graft_entry__.build()). siss_b200 has no CPU or PyTorch fallback.") from e
    try:
        lib = ctypes.CDLL(str(LIB_PATH))
    except OSError as e:  # pragma: no cover - depends on the environment
        raise SissLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise SissLibraryError(f"{LIB_PATH} does not export {name}; rebuild it") from e
        fn.restype = res
        fn.argtypes = args
    got = lib.siss_abi_version()
    if got != ABI_VERSION:
        raise SissLibraryError(f"{LIB_PATH} has ABI version {got}, python layer expects {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


_ext = None


def load_ext():
    """The thin torch extension over the same C ABI (csrc/torch_ext.cpp): returns ``torch.ops.siss_b200``. Built in
    place on demand (g++ against the torch headers) and checked for staleness by content hash, like the library."""
    global _ext
    if _ext is not None:
        return _ext
    import torch
    from . import build as _build
    load()
    if _build.ext_needs_build():
        try:
            _build.build_torch_ext()
        except Exception as e:
            raise SissLibraryError(f"{_build.EXT_PATH} is missing or stale and could not be built ({e}); set "
                                   "SISS_BINDING=ctypes to use the ctypes binding of the same library") from e
    try:
        torch.ops.load_library(str(_build.EXT_PATH))
    except OSError as e:  # pragma: no cover - depends on the environment
        raise SissLibraryError(f"cannot load {_build.EXT_PATH}: {e}") from e
    _ext = torch.ops.siss_b200
    return _ext


def error_string(code: int) -> str:
    return load().siss_error_string(code).decode()


def check(code: int, what: str) -> None:
    if code != 0:
        raise SissLibraryError(f"{what} failed with code {code}: {error_string(code)}")


_devices_checked = set()


def require_b200(device_index: int = -1) -> None:
    """Raise unless the CUDA device the op runs on is sm_100. Checked once per device (the library keeps its
    per-device launch state — shared-memory opt-ins, SM counts — keyed by device as well)."""
    if device_index in _devices_checked:
        return
    sm, major, minor = c_int(), c_int(), c_int()
    check(load().siss_check_device(ctypes.byref(sm), ctypes.byref(major), ctypes.byref(minor)), "siss_check_device")
    _devices_checked.add(device_index)
