"""Aspect-ratio bucketing and rank sharding -- host-side index logic of the path.

Public surface mirrors ``modules/dataset/bucket.py:32-214`` (``BucketManager``) and the two aspect samplers of
``modules/dataset/samplers.py:12-34,73-170``.  This is integer / PRNG work with no kernel; what matters is that every
rank draws EXACTLY the reference's id and resolution sequence for a given seed, so the implementation consumes the two
legacy ``numpy.random.RandomState`` streams call for call:

* ``bucket_prng = RandomState(seed)``; ``sharding_prng = RandomState(bucket_prng.tomaxint() % (2**32 - 1))``
* per epoch: one ``sharding_prng.shuffle`` of all ids -> truncate to a multiple of ``batch * world`` -> ``[rank::world]``;
  one ``bucket_prng.shuffle`` per non-empty bucket (grid order); remainders (``len % batch``) come off the FRONT
* per batch: one ``bucket_prng.choice(n, 1, p=float32 weights)`` while any bucket is left, plus a shuffle of the
  left-overs when they are chosen.

Reference quirks that change the sequence are kept and marked ``# ref quirk``: emptiness is tested with ``any(ids)``,
so id 0 counts as "nothing" (``bucket.py:56,135,145,192``).
"""
from __future__ import annotations

import copy
import random
from dataclasses import dataclass, field
from typing import Hashable, Iterator, Optional

import numpy as np

Size = tuple  # (width, height)
LEFT_OVER = "left_over"


def _truthy(ids) -> bool:
    """``any(ids)`` exactly as the reference evaluates it (id 0 is falsy)."""  # ref quirk
    return any(ids)


@dataclass
class Bucket:
    size: Size
    ids: list = field(default_factory=list)

    def __hash__(self) -> int:
        return hash(self.size)

    def __str__(self) -> str:
        return str(self.size)

    @property
    def aspect(self) -> float:
        return float(self.size[0]) / float(self.size[1])


def bucket_grid(base_res=(512, 512), max_size=768 * 512, dim_range=(256, 1024), divisor=64) -> list[Size]:
    """Sorted (w, h) grid of ``bucket.py:60-85``.  For every width the tallest height under the area cap, then the
    same with the roles swapped; the second sweep's outer test is ``h / min_dim <= max_size`` (sic, ``:75``) and is
    kept because it is part of the observable grid."""
    lo, hi = dim_range
    sizes = set()
    w = lo
    while w * lo <= max_size and w <= hi:
        h = lo
        while w * (h + divisor) <= max_size and (h + divisor) <= hi:
            if (w, h) == base_res:
                sizes.add(base_res)
            h += divisor
        sizes.add((w, h))
        w += divisor
    h = lo
    while h / lo <= max_size and h <= hi:  # ref quirk: ratio, not area
        w = lo
        while h * (w + divisor) <= max_size and (w + divisor) <= hi:
            w += divisor
        sizes.add((w, h))
        h += divisor
    return sorted(sizes)


class BucketManager:
    def __init__(self, batch_size: int, seed: Optional[int] = None, world_size=1, global_rank=0):
        self.batch_size, self.world_size, self.global_rank = batch_size, world_size, global_rank
        self.buckets: Optional[list[Bucket]] = None
        self.id_size_map: dict = {}
        self.base_res: Optional[Size] = None
        self.epoch: Optional[dict] = None
        self.epoch_remainders: Optional[list] = None
        self.batch_total = 0
        self.batch_delivered = 0
        self.bucket_prng = np.random.RandomState(seed)
        self.sharding_prng = np.random.RandomState(self.bucket_prng.tomaxint() % (2 ** 32 - 1))

    # ---- state predicates (bucket.py:52-58) -------------------------------------------------------
    @property
    def epoch_null(self) -> bool:
        return self.epoch is None or self.epoch_remainders is None

    @property
    def epoch_empty(self) -> bool:
        nothing_left = not (_truthy(self.epoch_remainders) or len(self.epoch) > 0)
        return nothing_left or self.batch_total == self.batch_delivered

    # ---- setup ---------------------------------------------------------------------------------------
    def gen_buckets(self, base_res=(512, 512), max_size=768 * 512, dim_range=(256, 1024), divisor=64) -> None:
        self.base_res = tuple(base_res)
        self.buckets = [Bucket(s) for s in bucket_grid(tuple(base_res), max_size, tuple(dim_range), divisor)]

    def put_in(self, id_size_map: dict, max_aspect_error=0.5) -> list:
        """Nearest-aspect assignment (``bucket.py:87-108``); first bucket wins ties; returns the skipped ids."""
        self.id_size_map = id_size_map
        skipped = []
        for ident, (w, h) in id_size_map.items():
            aspect = float(w) / float(h)
            best, best_err = None, None
            for b in self.buckets:
                err = abs(b.aspect - aspect)
                if best is None or err < best_err:
                    best, best_err = b, err
            if best_err < max_aspect_error:
                best.ids.append(ident)
            else:
                skipped.append(ident)
        return skipped

    # ---- per epoch --------------------------------------------------------------------------------------
    def local_ids(self) -> set:
        """This rank's share of the epoch (``bucket.py:110-124``)."""
        ids = list(self.id_size_map.keys())
        total = len(ids)
        self.sharding_prng.shuffle(ids)
        ids = ids[:total - (total % (self.batch_size * self.world_size))]
        ids = ids[self.global_rank::self.world_size]
        assert len(ids) % self.batch_size == 0
        self.batch_total = len(ids) // self.batch_size
        return set(ids)

    _get_local_ids = local_ids

    def start_epoch(self) -> None:
        mine = self.local_ids()
        epoch, remainders = {}, []
        for b in self.buckets:
            if not _truthy(b.ids):
                continue
            chosen = [i for i in b.ids if i in mine]
            self.bucket_prng.shuffle(chosen)
            extra = len(chosen) % self.batch_size
            if extra:
                remainders.extend(chosen[:extra])
                chosen = chosen[extra:]
            if not _truthy(chosen):
                continue
            epoch[b] = chosen
        self.epoch, self.epoch_remainders, self.batch_delivered = epoch, remainders, 0

    def get_batch(self):
        """One batch of ids and its (w, h) (``bucket.py:154-207``)."""
        if self.epoch_null:
            raise Exception("No epoch")
        bs = self.batch_size
        while True:
            keys = list(self.epoch.keys())
            weights = [len(self.epoch[k]) for k in keys]
            if len(self.epoch_remainders) >= bs:
                keys.append(LEFT_OVER)
                weights.append(len(self.epoch_remainders))
            probs = np.array(weights, dtype=np.float32)
            probs /= probs.sum()
            if len(self.epoch) > 0:
                pick = keys[int(self.bucket_prng.choice(len(keys), 1, p=probs)[0])]
            else:
                pick = LEFT_OVER
            if isinstance(pick, str):
                pool = self.epoch_remainders
                self.bucket_prng.shuffle(pool)
                batch, self.epoch_remainders = pool[:bs], pool[bs:]
                size = self.base_res
                break
            pool = self.epoch[pick]
            if len(pool) >= bs:
                batch, rest = pool[:bs], pool[bs:]
                self.epoch[pick] = rest
                size = pick.size
                if not _truthy(rest):
                    del self.epoch[pick]
                break
            # too few left for a batch: hand them to the left-overs and draw again
            self.epoch_remainders.extend(pool)
            del self.epoch[pick]
            assert len(self.epoch_remainders) >= bs or len(self.epoch) > 0
        self.batch_delivered += 1
        return batch, size

    def generator(self) -> Iterator:
        if self.epoch_null or self.epoch_empty:
            self.start_epoch()
        while not self.epoch_empty:
            yield self.get_batch()


# ---- samplers (modules/dataset/samplers.py) ---------------------------------------------------------------
@dataclass
class Index:
    """``modules/dataset/datasets.py:45-48``."""
    value: int
    size: Size


def scale_bucket_params(dim: int, c_size: float, c_dim: float, c_div: float) -> dict:
    return {"base_res": (dim, dim), "max_size": int(dim ** 2 * c_size), "dim_range": (int(dim / c_dim), int(dim * c_dim)),
            "divisor": int(dim / c_div)}


def get_gen_bucket_params(dim: int, bucket_config: dict) -> dict:
    params = scale_bucket_params(dim, bucket_config["c_size"], bucket_config["c_dim"], bucket_config["c_div"])
    manual = bucket_config.get("manual")
    if manual is not None:
        params = {**params, **dict(manual)}
    return params


DEFAULT_BUCKET_CONFIG = {"c_size": 1.5, "c_dim": 2.0, "c_div": 8.0, "max_aspect_error": 0.5}  # configs/__reserved_default__.yaml:52-57


class AspectSampler:
    """``samplers.py:73-105``: yields ``Index(id, (w, h))`` batch after batch for this rank."""

    def __init__(self, id_size_map: dict, base_size: int, bucket_config: dict, batch_size: int, seed: int, world_size=1,
                 global_rank=0):
        self.bucket_manager = BucketManager(batch_size, seed, world_size, global_rank)
        self.bucket_manager.gen_buckets(**get_gen_bucket_params(base_size, bucket_config))
        self.bucket_manager.put_in(id_size_map, bucket_config["max_aspect_error"])
        self._batch_size = batch_size

    def __iter__(self):
        for batch, size in self.bucket_manager.generator():
            for i in batch:
                yield Index(i, size)

    def batches(self):
        """Whole batches ``(ids, (w, h))`` -- what the train loop consumes."""
        yield from self.bucket_manager.generator()

    def __len__(self):
        if self.bucket_manager.epoch_null:
            self.bucket_manager.start_epoch()
        return self.bucket_manager.batch_total * self._batch_size


class AspectSamplerDB:
    """``samplers.py:108-170``: DreamBooth pairing -- every instance id is paired with a class id of the same bucket
    (nearest aspect when that bucket has no class image), drawn with the global ``random`` module like the reference."""

    def __init__(self, instance_id_size_map: dict, class_id_size_map: dict, base_size: int, bucket_config: dict,
                 batch_size: int, seed: int, world_size=1, global_rank=0):
        bm = BucketManager(batch_size, seed, world_size, global_rank)
        bm.gen_buckets(**get_gen_bucket_params(base_size, bucket_config))
        pristine = copy.deepcopy(bm.buckets)
        bm.put_in(instance_id_size_map, bucket_config["max_aspect_error"])
        self.bucket_manager = bm
        self._batch_size = batch_size
        cbm = BucketManager(1, seed, world_size, global_rank)
        cbm.buckets, cbm.base_res = pristine, bm.base_res
        cbm.put_in(class_id_size_map, bucket_config["max_aspect_error"])
        self.class_bucket_id_map: dict = {}
        for batch, size in cbm.generator():
            self.class_bucket_id_map.setdefault(size, []).append(batch[0])

    def _closest_class_entries(self, size):
        target = size[0] / size[1]
        best = min(self.class_bucket_id_map.keys(), key=lambda s: abs(s[0] / s[1] - target))
        return self.class_bucket_id_map[best]

    def __iter__(self):
        for batch, size in self.bucket_manager.generator():
            for instance_id in batch:
                pool = self.class_bucket_id_map.get(size)
                if not (pool is not None and _truthy(pool)):
                    pool = self._closest_class_entries(size)
                yield Index(instance_id, size), Index(random.choice(pool), size)

    def __len__(self):
        if self.bucket_manager.epoch_null:
            self.bucket_manager.start_epoch()
        return self.bucket_manager.batch_total * self._batch_size


def collate_order(pairs: list) -> list:
    """``modules/dataset/__init__.py:77-86``: instance items first, class items after -- the ordering
    ``torch.chunk(loss, 2)`` relies on."""
    if pairs and isinstance(pairs[0], tuple):
        return [a for a, _ in pairs] + [b for _, b in pairs]
    return list(pairs)
