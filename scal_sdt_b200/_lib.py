"""ctypes binding of ``libsdt_b200.so`` -- the C ABI declared in ``include/sdt_b200.h``.

Every entry point takes raw device pointers and a ``cudaStream_t``; PyTorch only owns the memory and
the stream.  There is no fallback: if the library is missing or the device is not sm_100, calls raise.
"""
from __future__ import annotations

import ctypes
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p
from pathlib import Path

from .build import LIB_PATH

SDT_F32, SDT_BF16, SDT_F16 = 0, 1, 2
TARGET_EPSILON, TARGET_SAMPLE, TARGET_V = 0, 1, 2


class SdtError(RuntimeError):
    pass


class PackSite(ctypes.Structure):
    _fields_ = [("A", c_void_p), ("B", c_void_p), ("A_p", c_void_p), ("At_p", c_void_p), ("B_p", c_void_p),
                ("Bt_p", c_void_p), ("K", c_int32), ("N", c_int32), ("r_true", c_int32), ("r", c_int32)]


class LoraProblem(ctypes.Structure):
    _fields_ = [("x", c_void_p), ("w", c_void_p), ("bias", c_void_p), ("A", c_void_p), ("B", c_void_p), ("y", c_void_p),
                ("t_save", c_void_p), ("residual", c_void_p)]


class LoraBwdProblem(ctypes.Structure):
    _fields_ = [("dy", c_void_p), ("x", c_void_p), ("wt", c_void_p), ("At", c_void_p), ("Bt", c_void_p), ("t_save", c_void_p),
                ("g_ws", c_void_p), ("dA", c_void_p), ("dB", c_void_p)]


class WgradSite(ctypes.Structure):
    _fields_ = [("x", c_void_p), ("g_ws", c_void_p), ("dA", c_void_p), ("dy", c_void_p), ("t_save", c_void_p), ("dB", c_void_p),
                ("M", c_int64), ("K", c_int64), ("N", c_int64)]


MAX_GROUP = 4
MAX_MULTI = 32


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
    "sdt_lora_linear_fwd_res": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p,
                                        c_int64, c_int64, c_int64, c_int, c_int, c_void_p]),
    "sdt_lora_linear_fwd_multi_supported": (c_int, [c_int, c_int64, c_int64, c_void_p, c_int]),
    "sdt_lora_linear_fwd_multi": (c_int, [c_void_p, c_void_p, c_int, c_float, c_int64, c_int64, c_int, c_int, c_void_p]),
    "sdt_lora_linear_geglu_supported": (c_int, [c_int64, c_int64, c_int64, c_int]),
    "sdt_lora_linear_geglu_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p,
                                          c_int64, c_int64, c_int64, c_int, c_int, c_void_p]),
    "sdt_lora_linear_fwd_group": (c_int, [c_void_p, c_int, c_float, c_int64, c_int64, c_int64, c_int, c_int, c_void_p]),
    "sdt_lora_linear_bwd_group_supported": (c_int, [c_int, c_int, c_int64, c_int64, c_int64, c_int]),
    "sdt_lora_linear_bwd_group": (c_int, [c_void_p, c_int, c_float, c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_int,
                                          c_void_p, c_void_p]),
    "sdt_lora_wgrad_workspace_bytes": (c_size_t, []),
    "sdt_lora_wgrad_max_sites": (c_int, []),
    "sdt_lora_wgrad_batch": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "sdt_lora_linear_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_int,
                                    c_void_p, c_void_p]),
    "sdt_lora_dropout": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_float, c_void_p, c_uint64, c_int, c_int, c_void_p]),
    "sdt_lora_pack": (c_int, [c_void_p, c_int, c_int64, c_int, c_void_p]),
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
    "sdt_group_norm_workspace_floats": (c_int64, [c_int64, c_int]),
    "sdt_group_norm_nhwc": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int,
                                    c_float, c_int, c_void_p, c_int64, c_void_p]),
    "sdt_group_norm_nhwc_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                        c_int64, c_int64, c_int, c_int, c_float, c_int, c_void_p]),
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
    if dtype == torch.float16:
        return SDT_F16
    raise SdtError(f"unsupported dtype {dtype}: libsdt_b200 computes in float32, bfloat16 or float16 (there is no fallback path)")


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


_wgrad_ws = {}


def wgrad_workspace() -> int:
    """Device pointer of the deterministic dA / dB reduction workspace of the CURRENT (device, stream): allocated and zeroed
    once, then reused by every backward launch on that stream (the kernels leave its counters at zero).  SDT_WGRAD_ATOMIC=1
    selects the atomic accumulation instead (A/B measurements; not reproducible run to run)."""
    import os

    import torch
    if os.environ.get("SDT_WGRAD_ATOMIC", "0") == "1":
        return 0
    st = torch.cuda.current_stream()
    key = (st.device.index, st.cuda_stream)
    ws = _wgrad_ws.get(key)
    if ws is None:
        ws = torch.zeros(load().sdt_lora_wgrad_workspace_bytes(), dtype=torch.uint8, device=st.device)
        _wgrad_ws[key] = ws
    return ws.data_ptr()


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def require_cuda(*tensors) -> None:
    """Every tensor handed to the library must live on the CURRENT CUDA device: the kernels are enqueued on the current
    device's stream (``stream_ptr``), one process per GPU.  A tensor on another GPU is an error, not a silent cross-device
    launch."""
    import torch
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise SdtError("scal_sdt_b200 runs on CUDA (sm_100a) tensors only; got a CPU tensor and there is no CPU path")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise SdtError(f"tensor lives on cuda:{t.device.index} but the current device is cuda:{cur}: call "
                           "torch.cuda.set_device first (one process per GPU; kernels go to the current device's stream)")


class PinnedRing:
    """A few pinned host slots for small per-step scalars that a captured graph reads from a fixed device buffer.

    ``push(values)`` writes the next slot and issues the async H2D copy into ``dev``; before a slot is rewritten the event
    recorded after its previous copy is waited for, so the host can run several steps ahead of the GPU without a DMA ever
    reading values of a later step."""

    def __init__(self, dev, slots: int = 4):
        import torch
        self.dev = dev
        self.host = [torch.zeros(dev.shape, dtype=dev.dtype).pin_memory() for _ in range(slots)]
        self.events = [None] * slots
        self.next = 0

    def push(self, fill) -> None:
        import torch
        i = self.next
        self.next = (i + 1) % len(self.host)
        if self.events[i] is not None:
            self.events[i].synchronize()
        fill(self.host[i])
        self.dev.copy_(self.host[i], non_blocking=True)
        ev = self.events[i] or torch.cuda.Event()
        ev.record()
        self.events[i] = ev


_checked_devices = set()


def device_check() -> None:
    import torch
    dev = torch.cuda.current_device()
    if dev not in _checked_devices:
        check(load().sdt_device_check(), "sdt_device_check")
        _checked_devices.add(dev)
