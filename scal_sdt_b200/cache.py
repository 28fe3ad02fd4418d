"""Latent / condition cache -- the step in front of the hot path.

The reference pre-encodes every image once (``cache.py:19-156``) into ONE safetensors file and trains from it
(``modules/dataset/datasets.py:70-91,147-152``); the VAE and the text encoder never run in the training loop.  This module
keeps that on-disk layout byte for byte, so caches written by the reference load here and vice versa:

    tensors   "{id}.latent.{k}"   [4, h, w]   k = 0 .. aug_group_size-1   (one encoded augmentation each)
              "{id}.cond"         [77, D]     (absent when the cache was built with --no-conds)
    metadata  {"json": json.dumps({"sizes": {"{id}.latent.{k}": [h, w]}, "entries": [ids], "total_entries": n,
                                   "aug_group_size": g})}                                   (``cache.py:129-154``)

and adds what a 30 ms training step needs in front of it: batches are collated straight into PINNED host buffers and copied
to the GPU on a side stream ``depth`` steps ahead, so the copy of step i+1 overlaps the kernels of step i
(``LatentDiffusionTrainer.graphed_step`` then copies device-to-device into the graph's static inputs).

Host code only -- no arithmetic of the path happens here.  Collation order is the reference's ``collate_fn``
(``modules/dataset/__init__.py:55-98``): instance items in sampler order, then all class items, which is what the
two-segment loss (``modules/model.py:338-342``) relies on.
"""
from __future__ import annotations

import json
import random
from collections import deque
from typing import Any, Iterable, Iterator, Mapping, Optional, Sequence

import torch

from .bucket import Size


def write_cache(path: str, latents: Mapping[int, Sequence[torch.Tensor]], conds: Optional[Mapping[int, torch.Tensor]] = None) -> dict:
    """Write a cache in the reference's layout from already encoded tensors.

    ``latents[id]`` is the list of encoded augmentations of entry ``id`` (all of one shape -- the reference asserts this,
    ``cache.py:142``); ``conds[id]`` its text condition.  Returns the metadata dict that was stored."""
    from safetensors.torch import save_file
    if not latents:
        raise ValueError("empty cache")
    groups = {len(v) for v in latents.values()}
    if len(groups) != 1:
        raise ValueError("every entry needs the same number of augmentations (aug_group_size)")
    aug = groups.pop()
    tensors: dict[str, torch.Tensor] = {}
    sizes: dict[str, list[int]] = {}
    for id_, group in latents.items():
        if len({tuple(t.shape) for t in group}) != 1:
            raise ValueError(f"entry {id_}: all augmentations must share one latent shape")
        for k, t in enumerate(group):
            key = f"{id_}.latent.{k}"
            tensors[key] = t.detach().cpu().contiguous()
            sizes[key] = list(t.shape[1:])               # [h, w] in latent pixels (cache.py:127)
        if conds is not None and id_ in conds:
            tensors[f"{id_}.cond"] = conds[id_].detach().cpu().contiguous()
    meta = {"sizes": sizes, "entries": list(latents.keys()), "total_entries": len(latents), "aug_group_size": aug}
    save_file(tensors, path, {"json": json.dumps(meta)})
    return meta


class LatentCache:
    """Reader for one cache file (``ImagePromptDataset`` with ``cache_file`` set, ``datasets.py:70-91``)."""

    def __init__(self, path: str):
        from safetensors import safe_open
        self.path = path
        self._f = safe_open(path, framework="pt", device="cpu")
        self.metadata: dict[str, Any] = json.loads(self._f.metadata()["json"])
        self.aug_group_size: int = int(self.metadata["aug_group_size"])
        self.entries: list[int] = list(self.metadata["entries"])
        self.has_conds = any(k.endswith(".cond") for k in self._f.keys())

    def __len__(self) -> int:
        return int(self.metadata["total_entries"])            # datasets.py:93-94

    def id_size_map(self) -> dict[int, Size]:
        """``AspectDataset.id_size_map`` for a cached dataset (``datasets.py:147-152``): ``Size(sizes["{id}.latent.0"])``.
        NOTE ``Size`` is ``tuple[int, int]`` read as (width, height) everywhere else, while the stored pair is the latent's
        ``[h, w]`` (``cache.py:127``): bucket assignment of a cached run sees latent height as width, in latent pixels.
        Reproduced as is -- a drop-in must bucket a cached dataset exactly like the reference does."""
        sizes = self.metadata["sizes"]
        return {k: tuple(sizes[f"{k}.latent.0"]) for k in self.entries}

    def latent(self, id_: int, aug_index: int) -> torch.Tensor:
        return self._f.get_tensor(f"{id_}.latent.{aug_index}")

    def cond(self, id_: int) -> torch.Tensor:
        return self._f.get_tensor(f"{id_}.cond")

    def item(self, id_: int, rng: Optional[random.Random] = None) -> tuple[torch.Tensor, Optional[torch.Tensor]]:
        """``CacheItem`` of ``__getitem__`` (``datasets.py:84-91``): a uniformly drawn augmentation and the condition."""
        r = rng if rng is not None else random
        k = r.randint(0, self.aug_group_size - 1)
        return self.latent(id_, k), (self.cond(id_) if self.has_conds else None)


def collate_ids(batch_ids: Iterable[Any]) -> list[tuple[str, int]]:
    """Order of ``collate_fn`` (``dataset/__init__.py:77-86``): plain ids / instance ids first, class ids after them.
    ``batch_ids`` holds ints (plain dataset) or ``(instance_id, class_id)`` pairs (DreamBooth).  Returns (role, id) pairs."""
    first: list[tuple[str, int]] = []
    classes: list[tuple[str, int]] = []
    for x in batch_ids:
        if isinstance(x, tuple):
            first.append(("instance", int(x[0])))
            classes.append(("class", int(x[1])))
        else:
            first.append(("instance", int(x)))
    return first + classes


class CachedBatchLoader:
    """Bucketed batches from cache files -> pinned host staging -> device, ``depth`` batches ahead of the consumer.

    ``batches`` yields ``(ids, size)`` as the samplers do (``AspectSampler`` / ``AspectSamplerDB`` / constant-size samplers):
    every latent of one batch has one shape.  ``class_cache`` supplies the class half of DreamBooth pairs.
    With ``device=None`` the loader yields the pinned (or plain, when CUDA is absent) host tensors themselves.
    """

    def __init__(self, cache: LatentCache, batches: Iterable[tuple[Sequence[Any], Any]], *, class_cache: Optional[LatentCache] = None,
                 device: Optional[torch.device] = None, depth: int = 2, seed: Optional[int] = None, dtype: Optional[torch.dtype] = None):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.cache, self.class_cache = cache, class_cache
        self.batches = batches
        self.device = torch.device(device) if device is not None else None
        self.depth = depth
        self.dtype = dtype
        self.rng = random.Random(seed)
        self._pin = torch.cuda.is_available()
        self._stream = torch.cuda.Stream(self.device) if self.device is not None and self.device.type == "cuda" else None
        # staging buffers are recycled per shape: (depth + 1) rotating slots so a slot is never rewritten while its copy may
        # still be in flight
        self._slots: dict[tuple, list[dict[str, torch.Tensor]]] = {}
        self._slot_next: dict[tuple, int] = {}
        self.bytes_staged = 0

    # ---- host side ------------------------------------------------------------------------------------------------
    def _staging(self, n: int, lat_shape: tuple, cond_shape: Optional[tuple], lat_dtype, cond_dtype) -> dict[str, torch.Tensor]:
        key = (n, lat_shape, cond_shape, lat_dtype, cond_dtype)
        ring = self._slots.setdefault(key, [])
        if len(ring) < self.depth + 1:
            slot = {"latents": torch.empty((n, *lat_shape), dtype=lat_dtype, pin_memory=self._pin)}
            if cond_shape is not None:
                slot["conds"] = torch.empty((n, *cond_shape), dtype=cond_dtype, pin_memory=self._pin)
            ring.append(slot)
            return slot
        i = self._slot_next.get(key, 0)
        self._slot_next[key] = (i + 1) % len(ring)
        slot = ring[i]
        ev = slot.get("_event")
        if ev is not None:
            ev.synchronize()                      # the H2D copy that last read this slot has finished
        return slot

    def _collate(self, ids: Sequence[Any]) -> dict[str, Any]:
        order = collate_ids(ids)
        items = []
        for role, id_ in order:
            src = self.class_cache if role == "class" else self.cache
            if src is None:
                raise ValueError("DreamBooth pairs need class_cache")
            items.append(src.item(id_, self.rng))
        lat0, cond0 = items[0]
        for lat, _ in items:
            if lat.shape != lat0.shape:
                raise ValueError(f"one batch must hold one latent shape (got {tuple(lat.shape)} and {tuple(lat0.shape)}): "
                                 "batches come from the aspect-ratio bucket sampler")
        has_cond = all(c is not None for _, c in items)
        lat_dtype = self.dtype or lat0.dtype
        cond_dtype = (self.dtype or cond0.dtype) if has_cond else None
        slot = self._staging(len(items), tuple(lat0.shape), tuple(cond0.shape) if has_cond else None, lat_dtype, cond_dtype)
        for i, (lat, cond) in enumerate(items):
            slot["latents"][i].copy_(lat)                     # stack + dtype cast straight into the pinned buffer
            if has_cond:
                slot["conds"][i].copy_(cond)
        self.bytes_staged += sum(v.numel() * v.element_size() for k, v in slot.items() if isinstance(v, torch.Tensor))
        out = {"ids": [id_ for _, id_ in order], "latents": slot["latents"]}
        if has_cond:
            out["conds"] = slot["conds"]
        out["_slot"] = slot
        return out

    # ---- device side ----------------------------------------------------------------------------------------------
    def _to_device(self, host: dict[str, Any]) -> dict[str, Any]:
        slot = host.pop("_slot")
        if self.device is None:
            # the consumer copies from the pinned slot itself (LatentDiffusionTrainer.graphed_step): it records an event after
            # its copy into slot["_event"], which _staging waits for before the slot is rewritten
            if self._pin:
                host["_host_slot"] = slot
            return host
        out = {"ids": host["ids"]}
        if self._stream is not None:
            with torch.cuda.stream(self._stream):
                for k in ("latents", "conds"):
                    if k in host:
                        out[k] = host[k].to(self.device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._stream)
            slot["_event"] = ev
            out["_ready"] = ev
        else:
            for k in ("latents", "conds"):
                if k in host:
                    out[k] = host[k].to(self.device)
        return out

    def __iter__(self) -> Iterator[dict[str, Any]]:
        queue: deque[dict[str, Any]] = deque()
        it = iter(self.batches)
        exhausted = False
        while True:
            while not exhausted and len(queue) < self.depth:
                try:
                    ids, _size = next(it)
                except StopIteration:
                    exhausted = True
                    break
                queue.append(self._to_device(self._collate(ids)))
            if not queue:
                return
            batch = queue.popleft()
            ev = batch.pop("_ready", None)
            if ev is not None:
                torch.cuda.current_stream(self.device).wait_event(ev)      # consumer stream waits, the host does not
                for k in ("latents", "conds"):
                    if k in batch:
                        batch[k].record_stream(torch.cuda.current_stream(self.device))
            yield batch
