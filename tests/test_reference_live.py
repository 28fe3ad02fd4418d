"""CPU tier, build container only: the product host code against the reference's OWN source executed live
(skipped where /root/reference does not exist, e.g. on the GPU box -- the golden fixtures cover the same ground there)."""
import random

import numpy as np
import pytest
import torch
from torch import nn

from oracle import ema_ref, reference_shim as shim
from scal_sdt_b200 import bucket as B

pytestmark = pytest.mark.skipif(not shim.available(), reason="/root/reference is not present")


def test_bucket_manager_live_sequences():
    ref = shim.load_reference_bucket()
    sizes = [(512, 512), (640, 448), (448, 640), (1024, 512), (300, 900), (768, 768), (832, 1216)]
    for seed, batch, world in ((1, 2, 1), (99, 4, 2), (114514, 3, 4)):
        rs = np.random.RandomState(seed)
        idmap = {i: sizes[int(k)] for i, k in enumerate(rs.randint(0, len(sizes), size=157))}
        for rank in range(world):
            a, b = ref.BucketManager(batch, seed, world, rank), B.BucketManager(batch, seed, world, rank)
            a.gen_buckets(); b.gen_buckets()
            a.put_in(dict(idmap), 0.5); b.put_in(dict(idmap), 0.5)
            for _ in range(3):
                ea = [([int(i) for i in ids], tuple(s)) for ids, s in a.generator()]
                eb = [([int(i) for i in ids], tuple(s)) for ids, s in b.generator()]
                assert ea == eb and a.batch_total == b.batch_total


def test_reference_ema_live_matches_restatement():
    ref = shim.load_reference_ema()
    torch.manual_seed(4)
    m1 = nn.Sequential(nn.Linear(9, 7), nn.Tanh(), nn.Linear(7, 3))
    m2 = nn.Sequential(nn.Linear(9, 7), nn.Tanh(), nn.Linear(7, 3))
    m2.load_state_dict(m1.state_dict())
    e1, e2 = ref.ExponentialMovingAverage(m1, 0.97), ema_ref.RefEMA(m2, 0.97)
    g = torch.Generator().manual_seed(0)
    for _ in range(25):
        with torch.no_grad():
            for p, q in zip(m1.parameters(), m2.parameters()):
                d = torch.randn(p.shape, generator=g)
                p.add_(d); q.add_(d)
        e1.update(); e2.update()
    for k, v in e1.shadow_params.items():
        assert torch.equal(v, e2.shadow_params[k])
    assert set(e1.state_dict()) == set(e2.state_dict())


def test_sampler_db_live():
    ref = shim.load_reference_samplers()
    rs = np.random.RandomState(3)
    sizes = [(512, 512), (768, 512), (512, 768), (640, 448)]
    inst = {i: sizes[int(k)] for i, k in enumerate(rs.randint(0, 4, size=40))}
    cls = {i: sizes[int(k)] for i, k in enumerate(rs.randint(0, 4, size=90))}

    class _Set:
        def __init__(self, m):
            self.id_size_map, self.image_paths = m, list(m)

    class _DB:
        def __init__(self, a, b):
            self.instance_set, self.class_set = _Set(a), _Set(b)

    cfg = ref.AttrDict(c_size=1.5, c_dim=2.0, c_div=8.0, max_aspect_error=0.5)
    random.seed(7)
    a = [(x.value, tuple(x.size), y.value) for x, y in ref.AspectSamplerDB(_DB(inst, cls), 512, cfg, 2, 5, 1, 0)]
    random.seed(7)
    b = [(x.value, tuple(x.size), y.value) for x, y in B.AspectSamplerDB(inst, cls, 512, B.DEFAULT_BUCKET_CONFIG, 2, 5, 1, 0)]
    assert a == b


def test_glue_goldens_regenerate_from_the_reference(tmp_path, monkeypatch):
    """``lora_glue.pt`` / ``denoise_steps.pt`` / ``config_module.json`` are what the reference's own source produces today."""
    import json

    from oracle import make_golden
    monkeypatch.setattr(make_golden, "GOLDEN", tmp_path)
    make_golden.make_lora_glue_fixture()
    make_golden.make_denoise_fixture()
    make_golden.make_config_module_fixture()
    real = make_golden.ROOT / "tests" / "golden"

    def same(a, b):
        if isinstance(a, torch.Tensor):
            return torch.equal(a, b)
        if isinstance(a, dict):
            return a.keys() == b.keys() and all(same(a[k], b[k]) for k in a)
        return a == b
    for name in ("lora_glue.pt", "denoise_steps.pt"):
        assert same(torch.load(tmp_path / name), torch.load(real / name)), name
    assert json.loads((tmp_path / "config_module.json").read_text()) == json.loads((real / "config_module.json").read_text())
