"""ctypes binding of ``libsdt_b200.so`` -- the C ABI declared in ``include/sdt_b200.h``.

Every entry point takes raw device pointers and a ``cudaStream_t``; PyTorch only owns the memory and
the stream.  There is no fallback: if the library is missing or the device is not sm_100, calls raise.
"""
from __future__ import annotations

import ctypes
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p
from pathlib import Path

from .build import LIB_PATH

SDT_F32, SDT_BF16 = 0, 1
TARGET_EPSILON, TARGET_SAMPLE, TARGET_V = 0, 1, 2


class SdtError(RuntimeError):
    pass


class PackSite(ctypes.Structure):
    _fields_ = [("A", c_void_p), ("B", c_void_p), ("A_p", c_void_p), ("At_p", c_void_p), ("B_p", c_void_p),
                ("Bt_p", c_void_p), ("K", c_int32), ("N", c_int32), ("r_true", c_int32), ("r", c_int32)]


class LoraProblem(ctypes.Structure):
    _fields_ = [("x", c_void_p), ("w", c_void_p), ("bias", c_void_p), ("A", c_void_p), ("B", c_void_p), ("y", c_void_p),
                ("t_save", c_void_p)]


class LoraBwdProblem(ctypes.Structure):
    _fields_ = [("dy", c_void_p), ("x", c_void_p), ("wt", c_void_p), ("At", c_void_p), ("Bt", c_void_p), ("t_save", c_void_p),
                ("g_ws", c_void_p), ("dA", c_void_p), ("dB", c_void_p)]


MAX_GROUP = 4


class Chunk(ctypes.Structure):
    _fields_ = [("tensor", c_int32), ("pad", c_int32), ("offset", c_int64)]


# name -> (restype, argtypes); mirrors include/sdt_b200.h one to one
SIGNATURES = {
    "sdt_version": (c_int, []),
    "sdt_last_error": (c_char_p, []),
    "sdt_device_check": (c_int, []),
    "sdt_launch_count": (ctypes.c_longlong, []),
    "sdt_lora_linear_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p,
                                    c_int64, c_int64, c_int64, c_int, c_int, c_void_p]),
    "sdt_lora_linear_fwd_group": (c_int, [c_void_p, c_int, c_float, c_int64, c_int64, c_int64, c_int, c_int, c_void_p]),
    "sdt_lora_linear_bwd_group_supported": (c_int, [c_int, c_int, c_int64, c_int64, c_int64, c_int]),
    "sdt_lora_linear_bwd_group": (c_int, [c_void_p, c_int, c_float, c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_int,
                                          c_void_p]),
    "sdt_lora_linear_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_int,
                                    c_void_p]),
    "sdt_lora_pack": (c_int, [c_void_p, c_int, c_int64, c_void_p]),
    "sdt_noise_target": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int64,
                                 c_int64, c_int, c_void_p, c_void_p]),
    "sdt_mse_loss_workspace_bytes": (c_size_t, []),
    "sdt_mse_loss": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                             c_int64, c_int64, c_float, c_float, c_void_p, c_void_p]),
    "sdt_ema_update_flat": (c_int, [c_void_p, c_void_p, c_int64, c_float, c_void_p, c_int, c_void_p]),
    "sdt_ema_update_multi": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_int,
                                     c_void_p]),
    "sdt_adamw_flat": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, ctypes.POINTER(c_float), c_void_p,
                               c_float, c_void_p, c_float, c_void_p, c_void_p]),
    "sdt_geglu": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p]),
    "sdt_group_norm_nhwc": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int,
                                    c_float, c_int, c_void_p]),
    "sdt_group_norm_nhwc_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                        c_int64, c_int, c_int, c_float, c_int, c_void_p]),
    "sdt_residual_bias_add": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "sdt_layer_norm_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float,
                                   c_void_p]),
    "sdt_layer_norm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "sdt_comm_unique_id": (c_int, [c_void_p]),
    "sdt_comm_init": (c_int, [c_void_p, c_int, c_int]),
    "sdt_comm_world": (c_int, []),
    "sdt_allreduce": (c_int, [c_void_p, c_int64, c_int, c_void_p]),
    "sdt_comm_destroy": (c_int, []),
    "sdt_simt_gemm_f32": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p,
                                  c_float, c_float, c_int64, c_int64, c_int64, c_void_p]),
    "sdt_debug_set": (c_int, [c_int, c_uint64]),
}

_lib = None


def library_path() -> Path:
    return LIB_PATH


def load() -> ctypes.CDLL:
    """Load the library (once).  Raises ``SdtError`` when it has not been built -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise SdtError(f"{LIB_PATH} is missing: run `python -m scal_sdt_b200.build` (or __graft_entry__.build()). "
                       "scal_sdt_b200 has no non-CUDA code path.")
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().sdt_last_error()
        raise SdtError(f"{what or 'libsdt_b200'} failed (code {rc}): {msg.decode() if msg else ''}")


def dtype_code(dtype) -> int:
    import torch
    if dtype == torch.float32:
        return SDT_F32
    if dtype == torch.bfloat16:
        return SDT_BF16
    raise SdtError(f"unsupported dtype {dtype}: libsdt_b200 computes in float32 or bfloat16 (fp16 is not implemented, "
                   "and there is no fallback path)")


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise SdtError("scal_sdt_b200 runs on CUDA (sm_100a) tensors only; got a CPU tensor and there is no CPU path")


_checked_devices = set()


def device_check() -> None:
    import torch
    dev = torch.cuda.current_device()
    if dev not in _checked_devices:
        check(load().sdt_device_check(), "sdt_device_check")
        _checked_devices.add(dev)
