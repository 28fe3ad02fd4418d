"""CPU tier: aspect-ratio bucketing / rank sharding / DreamBooth pairing are bit-exact with sequences produced by the
reference's own ``bucket.py`` / ``samplers.py`` (golden fixtures from ``oracle/make_golden.py``)."""
import json
import random
from pathlib import Path

import numpy as np
import pytest

from scal_sdt_b200 import bucket as B

GOLDEN = Path(__file__).parent / "golden"
SIZES8 = [(512, 512), (768, 512), (512, 768), (640, 448), (1024, 576), (576, 1024), (832, 1216), (900, 600)]


def id_size_map(n, seed=0):
    rs = np.random.RandomState(seed)
    return {i: tuple(SIZES8[int(k)]) for i, k in enumerate(rs.randint(0, len(SIZES8), size=n))}


def test_bucket_grids_match_reference():
    grids = json.loads((GOLDEN / "bucket_grids.json").read_text())
    for name, g in grids.items():
        p = g["params"]
        got = B.bucket_grid(tuple(p["base_res"]), p["max_size"], tuple(p["dim_range"]), p["divisor"])
        assert [list(s) for s in got] == g["sizes"], name
    assert len(grids["default512"]["sizes"]) == 19 and len(grids["manual786432"]["sizes"]) == 23
    assert [768, 1024] in grids["manual786432"]["sizes"] and [1024, 768] in grids["manual786432"]["sizes"]


def test_scaled_params():
    assert B.scale_bucket_params(512, 1.5, 2.0, 8.0) == {"base_res": (512, 512), "max_size": 393216,
                                                         "dim_range": (256, 1024), "divisor": 64}
    cfg = dict(B.DEFAULT_BUCKET_CONFIG, manual={"max_size": 786432})
    assert B.get_gen_bucket_params(512, cfg)["max_size"] == 786432


def test_epoch_sequences_bit_exact():
    grids = json.loads((GOLDEN / "bucket_grids.json").read_text())
    cases = json.loads((GOLDEN / "bucket_epochs.json").read_text())
    for name, c in cases.items():
        p = grids[c["grid"]]["params"]
        for rank_s, epochs in c["ranks"].items():
            bm = B.BucketManager(c["batch"], c["seed"], c["world"], int(rank_s))
            bm.gen_buckets(tuple(p["base_res"]), p["max_size"], tuple(p["dim_range"]), p["divisor"])
            bm.put_in(id_size_map(c["n_ids"], c["id_seed"]), 0.5)
            for ep in epochs:
                got = [[[int(i) for i in ids], list(size)] for ids, size in bm.generator()]
                assert bm.batch_total == ep["batch_total"], (name, rank_s)
                assert got == ep["batches"], (name, rank_s)


def test_batch_total_anchor():
    """seed 114514, batch 4, world 2, 200 ids -> 25 batches per rank."""
    c = json.loads((GOLDEN / "bucket_epochs.json").read_text())["s114514_b4_w2"]
    assert c["ranks"]["0"][0]["batch_total"] == 25 and c["ranks"]["1"][0]["batch_total"] == 25
    assert len(c["ranks"]["0"][0]["batches"]) == 25


def test_rank_shards_are_disjoint_and_equal_sized():
    maps = id_size_map(1000)
    seen = []
    for rank in range(8):
        bm = B.BucketManager(8, 114514, 8, rank)
        bm.gen_buckets()
        bm.put_in(maps, 0.5)
        ids = bm.local_ids()
        assert len(ids) == 120 and bm.batch_total == 15
        seen.append(ids)
    union = set().union(*seen)
    assert len(union) == sum(len(s) for s in seen) == 960


def test_sampler_db_pairs_bit_exact():
    gold = json.loads((GOLDEN / "sampler_db.json").read_text())
    for key, rec in gold["db"].items():
        world, rank = int(key[1]), int(key.split("_r")[1])
        random.seed(114514)
        s = B.AspectSamplerDB(id_size_map(120, 1), id_size_map(300, 2), 512, B.DEFAULT_BUCKET_CONFIG, 4, 114514, world, rank)
        assert len(s) == rec["len"]
        got = [[a.value, list(a.size), b.value, list(b.size)] for a, b in s]
        assert got == rec["pairs"], key
    s = B.AspectSampler(id_size_map(90, 3), 512, B.DEFAULT_BUCKET_CONFIG, 4, 42)
    assert len(s) == gold["plain"]["len"]
    assert [[a.value, list(a.size)] for a in s] == gold["plain"]["items"]


def test_collate_order_instance_then_class():
    pairs = [(("i", 0), ("c", 9)), (("i", 1), ("c", 8))]
    assert B.collate_order(pairs) == [("i", 0), ("i", 1), ("c", 9), ("c", 8)]
    assert B.collate_order([1, 2, 3]) == [1, 2, 3]


def test_empty_shard_yields_nothing():
    bm = B.BucketManager(4, 5, 3, 0)
    bm.gen_buckets()
    bm.put_in(id_size_map(10), 0.5)
    assert list(bm.generator()) == [] and bm.batch_total == 0
    with pytest.raises(Exception, match="No epoch"):
        B.BucketManager(1, 0).get_batch()
