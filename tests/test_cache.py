"""Latent / condition cache: on-disk layout of the reference's ``cache.py`` / ``datasets.py`` reader, collate
order of ``collate_fn``, and the pinned prefetching loader.  CPU tests; the device copy path has a ``gpu`` test."""
import json
import random

import pytest
import torch

from scal_sdt_b200.cache import CachedBatchLoader, LatentCache, collate_ids, write_cache


def make_cache(tmp_path, name, ids, aug=2, hw=(8, 6), seed=0, conds=True):
    g = torch.Generator().manual_seed(seed)
    lat = {i: [torch.randn(4, *hw, generator=g) for _ in range(aug)] for i in ids}
    cnd = {i: torch.randn(77, 16, generator=g) for i in ids} if conds else None
    path = str(tmp_path / name)
    meta = write_cache(path, lat, cnd)
    return path, lat, cnd, meta


def test_layout_is_the_reference_layout(tmp_path):
    """Keys / metadata exactly as ``cache.py:119-154`` writes them, read back the way ``datasets.py:70-91`` reads them."""
    from safetensors import safe_open
    path, lat, cnd, meta = make_cache(tmp_path, "c.safetensors", [3, 7, 11], aug=3)
    f = safe_open(path, framework="pt", device="cpu")            # the reference's reader, verbatim calls
    md = json.loads(f.metadata()["json"])
    assert set(md) == {"sizes", "entries", "total_entries", "aug_group_size"}
    assert md["aug_group_size"] == 3 and md["total_entries"] == 3 and md["entries"] == [3, 7, 11]
    assert md["sizes"]["7.latent.2"] == [8, 6]
    assert set(f.keys()) == {f"{i}.latent.{k}" for i in (3, 7, 11) for k in range(3)} | {f"{i}.cond" for i in (3, 7, 11)}
    assert torch.equal(f.get_tensor("11.latent.1"), lat[11][1]) and torch.equal(f.get_tensor("3.cond"), cnd[3])
    assert meta == md


def test_reader_matches_reference_semantics(tmp_path):
    path, lat, cnd, _ = make_cache(tmp_path, "c.safetensors", [0, 1, 2, 3], aug=4, hw=(12, 8))
    c = LatentCache(path)
    assert len(c) == 4 and c.aug_group_size == 4 and c.has_conds
    # id_size_map: Size(sizes["{id}.latent.0"]) -- the stored [h, w] pair as is (datasets.py:150-152)
    assert c.id_size_map() == {i: (12, 8) for i in range(4)}
    # augmentation draw: random.randint(0, aug_group_size - 1) per item (datasets.py:88-89)
    rng, ref = random.Random(5), random.Random(5)
    for i in (2, 0, 3, 3):
        latent, cond = c.item(i, rng)
        assert torch.equal(latent, lat[i][ref.randint(0, 3)]) and torch.equal(cond, cnd[i])


def test_collate_order_is_instances_then_classes():
    """``collate_fn`` (dataset/__init__.py:77-86): class items are appended after ALL instance items."""
    assert collate_ids([5, 2, 9]) == [("instance", 5), ("instance", 2), ("instance", 9)]
    assert collate_ids([(5, 50), (2, 20)]) == [("instance", 5), ("instance", 2), ("class", 50), ("class", 20)]


def test_loader_host_batches_and_dreambooth_halves(tmp_path):
    ipath, ilat, icnd, _ = make_cache(tmp_path, "inst.safetensors", [0, 1, 2, 3], aug=1, seed=1)
    cpath, clat, ccnd, _ = make_cache(tmp_path, "class.safetensors", [10, 11], aug=1, seed=2)
    batches = [([(0, 10), (3, 11)], (64, 48)), ([(1, 11), (2, 10)], (64, 48))]
    loader = CachedBatchLoader(LatentCache(ipath), batches, class_cache=LatentCache(cpath), device=None, depth=2, seed=0)
    out = list(loader)
    assert [b["ids"] for b in out] == [[0, 3, 10, 11], [1, 2, 11, 10]]
    b0 = out[0]
    assert b0["latents"].shape == (4, 4, 8, 6) and b0["conds"].shape == (4, 77, 16)
    # instance half first, class half second: what chunk(loss, 2) splits (model.py:338-342)
    assert torch.equal(b0["latents"][0], ilat[0][0]) and torch.equal(b0["latents"][1], ilat[3][0])
    assert torch.equal(b0["latents"][2], clat[10][0]) and torch.equal(b0["conds"][3], ccnd[11])


def test_loader_recycles_staging_without_clobbering(tmp_path):
    """More batches than staging slots: every yielded batch still holds its own data when it is consumed."""
    path, lat, cnd, _ = make_cache(tmp_path, "c.safetensors", list(range(12)), aug=1)
    batches = [([i, i + 1], (48, 64)) for i in range(0, 12, 2)]
    loader = CachedBatchLoader(LatentCache(path), batches, device=None, depth=2, seed=0, dtype=torch.bfloat16)
    for b, (ids, _) in zip(loader, batches):
        assert b["ids"] == list(ids) and b["latents"].dtype == torch.bfloat16
        assert torch.equal(b["latents"][0], lat[ids[0]][0].bfloat16()) and torch.equal(b["latents"][1], lat[ids[1]][0].bfloat16())
        assert torch.equal(b["conds"][1], cnd[ids[1]].bfloat16())


def test_loader_rejects_mixed_shapes_and_missing_class_cache(tmp_path):
    g = torch.Generator().manual_seed(0)
    path = str(tmp_path / "m.safetensors")
    write_cache(path, {0: [torch.randn(4, 8, 8, generator=g)], 1: [torch.randn(4, 8, 6, generator=g)]})
    c = LatentCache(path)
    assert not c.has_conds
    with pytest.raises(ValueError, match="one latent shape"):
        list(CachedBatchLoader(c, [([0, 1], (64, 64))], device=None))
    with pytest.raises(ValueError, match="class_cache"):
        list(CachedBatchLoader(c, [([(0, 1)], (64, 64))], device=None))
    with pytest.raises(ValueError, match="aug_group_size"):
        write_cache(path, {0: [torch.zeros(4, 2, 2)], 1: [torch.zeros(4, 2, 2), torch.zeros(4, 2, 2)]})


def test_loader_feeds_the_bucket_sampler(tmp_path):
    """End to end on the host: cache -> id_size_map -> AspectSampler batches -> loader."""
    from scal_sdt_b200.bucket import AspectSampler
    g = torch.Generator().manual_seed(3)
    shapes = [(8, 8), (12, 8), (8, 12)]
    lat = {i: [torch.randn(4, *shapes[i % 3], generator=g)] for i in range(24)}
    path = str(tmp_path / "arb.safetensors")
    write_cache(path, lat, {i: torch.randn(77, 8, generator=g) for i in range(24)})
    cache = LatentCache(path)

    cfg = {"c_size": 1.5, "c_dim": 2.0, "c_div": 2.0, "max_aspect_error": 0.5}       # latent-pixel sizes: base 8, divisor 4
    sampler = AspectSampler(cache.id_size_map(), 8, cfg, 2, 7)
    n = 0
    for b in CachedBatchLoader(cache, sampler.batches(), device=None, seed=1):
        assert b["latents"].shape[0] == 2 and len({tuple(lat[i][0].shape) for i in b["ids"]}) == 1
        n += 1
    assert n >= 6


@pytest.mark.gpu
def test_loader_prefetches_to_device(tmp_path):
    path, lat, cnd, _ = make_cache(tmp_path, "c.safetensors", list(range(8)), aug=1)
    batches = [([i, i + 1], (48, 64)) for i in range(0, 8, 2)]
    loader = CachedBatchLoader(LatentCache(path), batches, device=torch.device("cuda:0"), depth=2, seed=0)
    for b, (ids, _) in zip(loader, batches):
        assert b["latents"].is_cuda and b["conds"].is_cuda
        assert torch.equal(b["latents"].cpu(), torch.stack([lat[i][0] for i in ids]))
        assert torch.equal(b["conds"].cpu(), torch.stack([cnd[i] for i in ids]))
    assert loader.bytes_staged > 0
