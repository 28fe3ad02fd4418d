"""Generate ``tests/golden/`` by EXECUTING THE REFERENCE'S OWN SOURCE (``/root/reference``) -- test infrastructure.

Run in the build container only:  ``python -m oracle.make_golden``.  The outputs are small JSON / .pt fixtures that
travel with the repo; the GPU box never reads ``/root/reference``.

* ``bucket_grids.json``    ``BucketManager.gen_buckets`` (``modules/dataset/bucket.py:60-85``) for several parameter sets
* ``bucket_epochs.json``   id / resolution sequences of ``BucketManager.generator`` per (seed, batch, world, rank), two epochs
* ``sampler_db.json``      ``AspectSamplerDB`` instance/class pairing (``modules/dataset/samplers.py:108-170``)
* ``walker.json``          modules visited by ``apply_module_config`` (``modules/utils/torch/module.py:14-63``) for every
                           ``configs/optim_targets/*.yaml`` over the UNet skeleton (+ a CLIP-named text-encoder skeleton)
* ``ema_reference.pt``     shadow parameters produced by the reference ``modules/ema.py`` on an all-trainable module
"""
from __future__ import annotations

import json
import random
import sys
from pathlib import Path

import numpy as np
import torch
import yaml
from torch import nn

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import reference_shim as shim  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"

SIZES8 = [(512, 512), (768, 512), (512, 768), (640, 448), (1024, 576), (576, 1024), (832, 1216), (900, 600)]


def synthetic_id_size_map(n: int, seed: int = 0, sizes=SIZES8) -> dict:
    rs = np.random.RandomState(seed)
    return {i: tuple(sizes[int(k)]) for i, k in enumerate(rs.randint(0, len(sizes), size=n))}


GRID_PARAMS = {
    "default512": dict(base_res=(512, 512), max_size=768 * 512, dim_range=(256, 1024), divisor=64),
    "manual786432": dict(base_res=(512, 512), max_size=786432, dim_range=(256, 1024), divisor=64),
    "scaled768": dict(base_res=(768, 768), max_size=int(768 ** 2 * 1.5), dim_range=(384, 1536), divisor=96),
    "scaled640": dict(base_res=(640, 640), max_size=int(640 ** 2 * 1.5), dim_range=(320, 1280), divisor=80),
}

EPOCH_CASES = [
    # name, n_ids, seed, batch, world, grid
    ("s114514_b4_w2", 200, 114514, 4, 2, "default512"),
    ("s114514_b8_w8", 1000, 114514, 8, 8, "manual786432"),
    ("s7_b3_w1", 50, 7, 3, 1, "default512"),
    ("s0_b1_w1", 17, 0, 1, 1, "default512"),
    ("s5_b4_w3", 10, 5, 4, 3, "default512"),     # fewer ids than batch*world: empty shard
]


def make_bucket_fixtures():
    ref = shim.load_reference_bucket()
    grids = {}
    for name, params in GRID_PARAMS.items():
        bm = ref.BucketManager(1, 0)
        bm.gen_buckets(**params)
        grids[name] = {"params": {k: list(v) if isinstance(v, tuple) else v for k, v in params.items()},
                       "sizes": [list(b.size) for b in bm.buckets]}
    (GOLDEN / "bucket_grids.json").write_text(json.dumps(grids, indent=1))

    epochs = {}
    for name, n, seed, batch, world, grid in EPOCH_CASES:
        ranks = {}
        for rank in range(world):
            bm = ref.BucketManager(batch, seed, world, rank)
            bm.gen_buckets(**GRID_PARAMS[grid])
            bm.put_in(synthetic_id_size_map(n), 0.5)
            seq = []
            for _ in range(2):          # two epochs: the PRNG streams carry over
                ep = [[[int(i) for i in ids], list(size)] for ids, size in bm.generator()]
                seq.append({"batch_total": bm.batch_total, "batches": ep})
            ranks[str(rank)] = seq
        epochs[name] = {"n_ids": n, "seed": seed, "batch": batch, "world": world, "grid": grid, "id_seed": 0, "ranks": ranks}
    (GOLDEN / "bucket_epochs.json").write_text(json.dumps(epochs))


def make_sampler_fixture():
    ref = shim.load_reference_samplers()
    cfg = ref.AttrDict(c_size=1.5, c_dim=2.0, c_div=8.0, max_aspect_error=0.5)

    class _Set:
        def __init__(self, m):
            self.id_size_map = m
            self.image_paths = list(m)

    class _DB:
        def __init__(self, a, b):
            self.instance_set, self.class_set = _Set(a), _Set(b)

    out = {}
    for world in (1, 2):
        for rank in range(world):
            inst = synthetic_id_size_map(120, seed=1)
            cls = synthetic_id_size_map(300, seed=2)
            random.seed(114514)
            s = ref.AspectSamplerDB(_DB(inst, cls), 512, cfg, 4, 114514, world, rank)
            pairs = [[int(a.value), list(a.size), int(b.value), list(b.size)] for a, b in s]
            out[f"w{world}_r{rank}"] = {"len": len(s), "pairs": pairs}
    plain = {}
    s = ref.AspectSampler(_Set(synthetic_id_size_map(90, seed=3)), 512, cfg, 4, 42, 1, 0)
    plain["len"] = len(s)
    plain["items"] = [[int(a.value), list(a.size)] for a in s]
    (GOLDEN / "sampler_db.json").write_text(json.dumps({"db": out, "plain": plain}))


def clip_text_skeleton(layers=2, dim=16):
    """Module names of transformers' CLIPTextModel as far as the optim_targets YAMLs address them."""
    def layer():
        m = nn.Module()
        m.self_attn = nn.Module()
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            setattr(m.self_attn, n, nn.Linear(dim, dim))
        m.mlp = nn.Module()
        m.mlp.fc1, m.mlp.fc2 = nn.Linear(dim, 4 * dim), nn.Linear(4 * dim, dim)
        m.layer_norm1, m.layer_norm2 = nn.LayerNorm(dim), nn.LayerNorm(dim)
        return m
    te = nn.Module()
    te.text_model = nn.Module()
    te.text_model.encoder = nn.Module()
    te.text_model.encoder.layers = nn.ModuleList([layer() for _ in range(layers)])
    te.text_model.final_layer_norm = nn.LayerNorm(dim)
    return te


def make_walker_fixture():
    from scal_sdt_b200.unet import UNet2DConditionModel, UNetConfig
    walker = shim.load_reference_module_walker()
    out = {}
    tdir = shim.REFERENCE_ROOT / "configs" / "optim_targets"
    for path in sorted(tdir.glob("*.yaml")):
        from scal_sdt_b200.config import load_yaml
        cfg = load_yaml(path)
        entry = {"config": cfg, "visits": {}}
        for comp, module in (("unet", UNet2DConditionModel(UNetConfig.tiny())), ("text_encoder", clip_text_skeleton())):
            if cfg.get(comp) is None:
                continue
            visits = []

            def fn(sub, conf, mpath, _v=visits):
                _v.append([mpath, type(sub).__name__, {k: v for k, v in dict(conf).items() if k not in ("targets", "index")}])
            walker.apply_module_config(module, cfg[comp]["targets"], fn)
            entry["visits"][comp] = visits
        out[path.stem] = entry
    (GOLDEN / "walker.json").write_text(json.dumps(out))


def make_ema_fixture():
    ref = shim.load_reference_ema()
    torch.manual_seed(0)
    m = nn.Sequential(nn.Linear(33, 65), nn.GELU(), nn.Linear(65, 17), nn.LayerNorm(17))
    init = {k: v.clone() for k, v in m.state_dict().items()}
    ema = ref.ExponentialMovingAverage(m, 0.995)
    g = torch.Generator().manual_seed(3)
    deltas, decays = [], []
    for _ in range(12):
        step = []
        with torch.no_grad():
            for p in m.parameters():
                d = torch.randn(p.shape, generator=g) * 0.1
                p.add_(d)
                step.append(d)
        deltas.append(step)
        ema.update()
        decays.append(min(ema.decay, (1 + ema.num_updates) / (10 + ema.num_updates)))
    torch.save({"init": init, "deltas": deltas, "decays": decays, "num_updates": ema.num_updates,
                "shadow": {k: v.clone() for k, v in ema.shadow_params.items()}, "decay": 0.995}, GOLDEN / "ema_reference.pt")
    # the reference class on a partially frozen module (documents SURVEY fact 6)
    m2 = nn.Sequential(nn.Linear(4, 4), nn.Linear(4, 4))
    m2[0].weight.requires_grad_(False)
    try:
        e2 = ref.ExponentialMovingAverage(m2, 0.9)
        e2.update()
        crashed = None
    except KeyError as e:  # noqa: BLE001
        crashed = repr(e)
    (GOLDEN / "ema_reference_partial_freeze.json").write_text(json.dumps({"reference_raises": crashed}))


def main():
    if not shim.available():
        raise SystemExit("/root/reference is not present: golden fixtures can only be generated in the build container")
    GOLDEN.mkdir(parents=True, exist_ok=True)
    make_bucket_fixtures()
    make_sampler_fixture()
    make_walker_fixture()
    make_ema_fixture()
    for f in sorted(GOLDEN.iterdir()):
        print(f"{f.name}: {f.stat().st_size} bytes")


if __name__ == "__main__":
    main()
