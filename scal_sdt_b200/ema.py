"""Exponential moving average of the trainable parameters -- API of the reference's ``modules/ema.py:9-140``.

Same eight methods, same ``state_dict`` keys (``decay``, ``num_updates``, ``shadow_params`` keyed by parameter name
relative to the module).  Differences, both deliberate:

* ``update`` / ``apply`` / ``average_parameters`` walk the SHADOW keys.  The reference stores shadows only for
  ``requires_grad`` parameters (``ema.py:33-37``) but walks every ``named_parameters()`` entry (``ema.py:56-57``), so
  it raises ``KeyError`` on any partially frozen module, i.e. always under LoRA.  For all-trainable modules both walks
  visit the same tensors.
* the shadow set stays resident in HBM (the reference round-trips it CPU->GPU->CPU every step,
  ``modules/model.py:407-412``) and the whole update is ONE launch: ``s -= (1-d)(s-p)`` with the reference's three
  roundings (bit-exact in fp32), over a flat arena when the parameters are views of one, else multi-tensor.
"""
from __future__ import annotations

import contextlib
import copy
import ctypes

import torch
from torch import nn

from . import _lib
from ._lib import Chunk

_CHUNK_ELEMS = 1 << 16


class ExponentialMovingAverage:
    def __init__(self, module: nn.Module, decay: float, use_num_updates: bool = True):
        if decay < 0.0 or decay > 1.0:
            raise ValueError("Decay must be between 0 and 1")
        self.decay = decay
        self.num_updates = 0 if use_num_updates else None
        self.module = module
        tracked = [(name, p) for name, p in module.named_parameters() if p.requires_grad]
        self._names = [n for n, _ in tracked]
        self._flat = self._try_flat(tracked)
        if self._flat is not None:
            base, numel, offsets = self._flat
            self._shadow_flat = base.clone()
            self.shadow_params = {n: self._shadow_flat[o:o + p.numel()].view(p.shape) for (n, p), o in zip(tracked, offsets)}
        else:
            self._shadow_flat = None
            self.shadow_params = {n: p.clone().detach() for n, p in tracked}
        self._tables = None

    # ---- layout discovery ---------------------------------------------------------------------------
    @staticmethod
    def _try_flat(tracked):
        """If every tracked parameter is a view into one contiguous fp32 buffer (a ParamArena), return
        (buffer covering them, numel, per-parameter offsets)."""
        if not tracked or not all(p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() for _, p in tracked):
            return None
        st = tracked[0][1].untyped_storage()
        if any(p.untyped_storage().data_ptr() != st.data_ptr() for _, p in tracked):
            return None
        offs = [p.storage_offset() for _, p in tracked]
        lo = min(offs)
        hi = max(o + p.numel() for o, (_, p) in zip(offs, tracked))
        covered = sum(p.numel() for _, p in tracked)
        if hi - lo > covered + 4 * len(tracked):      # holes larger than alignment padding: not an arena
            return None
        p0 = tracked[0][1]
        base = torch.empty(0, dtype=torch.float32, device=p0.device).set_(st, lo, (hi - lo,), (1,))
        return base, hi - lo, [o - lo for o in offs]

    def _tracked(self):
        for name, p in self.module.named_parameters():
            if name in self.shadow_params:
                yield name, p

    @staticmethod
    def decay_at(decay: float, num_updates: int) -> float:
        """``ema.py:47-54`` after ``num_updates += 1``."""
        return min(decay, (1 + num_updates) / (10 + num_updates))

    def current_one_minus_decay(self) -> float:
        decay = self.decay
        if self.num_updates is not None:
            decay = self.decay_at(decay, self.num_updates)
        return 1.0 - decay

    def _build_tables(self, pairs):
        dev = pairs[0][1].device
        sp = torch.tensor([s.data_ptr() for s, _ in pairs], dtype=torch.int64).to(dev)
        pp = torch.tensor([p.data_ptr() for _, p in pairs], dtype=torch.int64).to(dev)
        ne = torch.tensor([p.numel() for _, p in pairs], dtype=torch.int64).to(dev)
        chunks = []
        for i, (_, p) in enumerate(pairs):
            for off in range(0, p.numel(), _CHUNK_ELEMS):
                chunks.append(Chunk(i, 0, off))
        arr = (Chunk * len(chunks))(*chunks)
        ch = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        key = tuple((s.data_ptr(), p.data_ptr()) for s, p in pairs)
        return key, sp, pp, ne, ch, len(chunks)

    @torch.no_grad()
    def update(self):
        """``ema.py:39-61``."""
        if self.num_updates is not None:
            self.num_updates += 1
        omd = self.current_one_minus_decay()
        lib = _lib.load()
        if self._flat is not None and self._shadow_flat is not None and self._shadow_flat.is_cuda:
            base, numel, _ = self._flat
            _lib.check(lib.sdt_ema_update_flat(self._shadow_flat.data_ptr(), base.data_ptr(), numel, omd, None,
                                               _lib.SDT_F32, _lib.stream_ptr()), "sdt_ema_update_flat")
            return
        pairs = [(self.shadow_params[n], p.detach()) for n, p in self._tracked()]
        if not pairs:
            return
        _lib.require_cuda(*[s for s, _ in pairs], *[p for _, p in pairs])
        by_dtype = {}
        for s, p in pairs:
            if s.dtype != p.dtype or not s.is_contiguous() or not p.is_contiguous():
                raise _lib.SdtError("EMA shadow / parameter must be contiguous and of equal dtype")
            by_dtype.setdefault(p.dtype, []).append((s, p))
        for dtype, group in by_dtype.items():
            key = tuple((s.data_ptr(), p.data_ptr()) for s, p in group)
            if self._tables is None or self._tables.get(dtype, (None,))[0] != key:
                self._tables = self._tables or {}
                self._tables[dtype] = self._build_tables(group)
            _, sp, pp, ne, ch, n_chunks = self._tables[dtype]
            _lib.check(lib.sdt_ema_update_multi(sp.data_ptr(), pp.data_ptr(), ne.data_ptr(), ch.data_ptr(), n_chunks,
                                                _CHUNK_ELEMS, omd, None, _lib.dtype_code(dtype), _lib.stream_ptr()),
                       "sdt_ema_update_multi")

    def _masters_changed(self) -> None:
        """The fp32 masters were written behind the optimizer's back: refresh the bf16 operands the kernels read."""
        from .lora import refresh_packed_operands
        refresh_packed_operands(self.module)

    def apply(self):
        """``ema.py:63-69``."""
        for name, p in self._tracked():
            p.data.copy_(self.shadow_params[name].data)
        self._masters_changed()

    @contextlib.contextmanager
    def average_parameters(self):
        """``ema.py:71-85``."""
        saved = {name: p.clone() for name, p in self._tracked()}
        self.apply()
        try:
            yield
        finally:
            for name, p in self._tracked():
                p.data.copy_(saved[name].data)
            self._masters_changed()

    def to(self, device=None, dtype=None) -> None:
        """``ema.py:87-99``.  Kept for API parity; the update itself needs the shadows on the GPU."""
        if self._shadow_flat is not None and dtype is None:
            self._shadow_flat = self._shadow_flat.to(device=device)
            _, _, offsets = self._flat
            self.shadow_params = {n: self._shadow_flat[o:o + s.numel()].view(s.shape)
                                  for (n, s), o in zip(self.shadow_params.items(), offsets)}
        else:
            self._shadow_flat = None
            self._flat = None
            self.shadow_params = {
                n: (p.to(device=device, dtype=dtype) if p.is_floating_point() else p.to(device=device))
                for n, p in self.shadow_params.items()}
        self._tables = None

    def state_dict(self) -> dict:
        """``ema.py:101-110``."""
        return {"decay": self.decay, "num_updates": self.num_updates, "shadow_params": self.shadow_params}

    def load_state_dict(self, state_dict: dict) -> None:
        """``ema.py:112-140``."""
        state_dict = copy.deepcopy(state_dict)
        self.decay = state_dict["decay"]
        if self.decay < 0.0 or self.decay > 1.0:
            raise ValueError("Decay must be between 0 and 1")
        self.num_updates = state_dict["num_updates"]
        assert self.num_updates is None or isinstance(self.num_updates, int), "Invalid num_updates"
        shadow = state_dict["shadow_params"]
        assert isinstance(shadow, dict), "shadow_params must be a dict"
        assert all(isinstance(p, torch.Tensor) for p in shadow.values()), "shadow_params must all be Tensors"
        if self._shadow_flat is not None and set(shadow) == set(self.shadow_params):
            for n, s in self.shadow_params.items():
                s.copy_(shadow[n])
        else:
            self._shadow_flat, self._flat = None, None
            self.shadow_params = shadow
        self._tables = None
