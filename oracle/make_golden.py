"""Generate ``tests/golden/`` by EXECUTING THE REFERENCE'S OWN SOURCE (``/root/reference``) -- test infrastructure.

Run in the build container only:  ``python -m oracle.make_golden``.  The outputs are small JSON / .pt fixtures that
travel with the repo; the GPU box never reads ``/root/reference``.

* ``bucket_grids.json``    ``BucketManager.gen_buckets`` (``modules/dataset/bucket.py:60-85``) for several parameter sets
* ``bucket_epochs.json``   id / resolution sequences of ``BucketManager.generator`` per (seed, batch, world, rank), two epochs
* ``sampler_db.json``      ``AspectSamplerDB`` instance/class pairing (``modules/dataset/samplers.py:108-170``)
* ``walker.json``          modules visited by ``apply_module_config`` (``modules/utils/torch/module.py:14-63``) for every
                           ``configs/optim_targets/*.yaml`` over the UNet skeleton (+ a CLIP-named text-encoder skeleton)
* ``ema_reference.pt``     shadow parameters produced by the reference ``modules/ema.py`` on an all-trainable module
* ``lora_glue.pt``         the reference's own ``modules/lora.py::get_lora`` (loralib stand-in = the restated layers): attribute /
                           state-dict / aliasing contract, forward + gradients on seeded inputs, the error on other module types
* ``denoise_steps.pt``     the reference's own ``LatentDiffusionModel._denoise_loss`` / ``training_step``
                           (``modules/model.py:289-348``) executed on a toy UNet for every prediction type, with and without
                           prior preservation: the draws it made, what it fed the UNet, per-element loss, reduced loss
* ``config_module.json``   the reference's own ``config_module`` (``modules/model.py:136-164``) on the UNet skeleton for every
                           optim_target YAML: param groups (names, optimizer overrides), trainable set, checkpoint keys written
                           by its ``on_save_checkpoint``; and ``get_optimizer``'s LR / weight-decay scaling
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


def make_lora_glue_fixture():
    ref = shim.load_reference_lora()
    cases = {}
    specs = {"linear_bias_r2_a4": (lambda: nn.Linear(6, 5), 2, 4, (3, 7, 6)),
             "linear_nobias_r4_a1": (lambda: nn.Linear(8, 3, bias=False), 4, 1, (5, 8)),
             "conv1x1_r2_a2": (lambda: nn.Conv2d(4, 6, 1), 2, 2, (2, 4, 3, 3))}
    for name, (make, rank, alpha, xshape) in specs.items():
        torch.manual_seed(11)
        base = make()
        base.requires_grad_(False)                     # config_module froze the whole network first (model.py:137)
        lora = ref.get_lora(base, rank, alpha)
        g = torch.Generator().manual_seed(12)
        with torch.no_grad():
            lora.lora_A.copy_(torch.randn(lora.lora_A.shape, generator=g) * 0.3)
            lora.lora_B.copy_(torch.randn(lora.lora_B.shape, generator=g) * 0.3)
        x = torch.randn(*xshape, generator=g, requires_grad=True)
        y = lora(x)
        dy = torch.randn(y.shape, generator=g)
        y.backward(dy)
        cases[name] = {
            "rank": rank, "alpha": alpha,
            "base_state": {k: v.clone() for k, v in base.state_dict().items()},
            "state_keys": sorted(lora.state_dict().keys()),
            "state_dtypes": {k: str(v.dtype) for k, v in lora.state_dict().items()},
            "buffers": sorted(n for n, _ in lora.named_buffers()),
            "requires_grad": {n: bool(p.requires_grad) for n, p in lora.named_parameters()},
            "weight_is_aliased": lora.weight is base.weight, "bias_is_aliased": lora.bias is base.bias,
            "scaling": float(lora.scaling), "lora_alpha": lora.lora_alpha.clone(),
            "has_python_lora_alpha_attr": "lora_alpha" in lora.__dict__,
            "lora_A": lora.lora_A.detach().clone(), "lora_B": lora.lora_B.detach().clone(),
            "x": x.detach().clone(), "dy": dy, "y": y.detach().clone(), "dx": x.grad.clone(),
            "dA": lora.lora_A.grad.clone(), "dB": lora.lora_B.grad.clone(),
            "frozen_grads_none": lora.weight.grad is None,
        }
    torch.manual_seed(0)
    fresh = ref.get_lora(nn.Linear(64, 32), 4, 1)
    cases["init"] = {"lora_B_all_zero": bool(torch.count_nonzero(fresh.lora_B) == 0),
                     "lora_A_absmax": float(fresh.lora_A.detach().abs().max()), "bound": 1 / 8}
    try:
        ref.get_lora(nn.LayerNorm(8))
        cases["other_module_error"] = None
    except Exception as e:  # noqa: BLE001
        cases["other_module_error"] = str(e)
    torch.save(cases, GOLDEN / "lora_glue.pt")


def make_denoise_fixture():
    from types import SimpleNamespace

    from oracle.diffusion_ref import RefDDIMScheduler
    model = shim.load_reference_model()

    class ToyUNet(nn.Module):
        """Stands where diffusers' UNet2DConditionModel stands: ``unet(noisy, t, conds).sample``; records what it was fed."""

        def __init__(self):
            super().__init__()
            self.mix = nn.Conv2d(4, 4, 1)
            self.seen = None

        dtype = torch.float32

        @property
        def device(self):
            return self.mix.weight.device

        def forward(self, noisy, timesteps, conds):
            self.seen = (noisy.detach().clone(), timesteps.detach().clone(), conds.detach().clone())
            shift = conds.mean(dim=(1, 2))[:, None, None, None]
            return SimpleNamespace(sample=self.mix(noisy) * (1 + 1e-3 * timesteps[:, None, None, None].float()) + shift)

    out = {}
    g = torch.Generator().manual_seed(21)
    for ptype in ("epsilon", "sample", "v"):
        for prior in (None, 0.6):
            B = 4
            torch.manual_seed(5)
            unet = ToyUNet()
            latents = torch.randn(B, 4, 6, 10, generator=g)
            conds = torch.randn(B, 7, 12, generator=g)
            cfg = shim.AttrDict({"prior_preservation": {"enabled": prior is not None, "prior_loss_weight": prior or 1.0}})
            me = shim.bare_lightning_module(model, unet=unet, scheduler=RefDDIMScheduler(ptype), config=cfg)
            seed = 1000 + len(out)
            torch.manual_seed(seed)
            loss_elem = me._denoise_loss(latents, conds)
            noisy, t_seen, conds_seen = unet.seen
            torch.manual_seed(seed)             # the draws of model.py:294,297-298, in the reference's order
            noise = torch.randn_like(latents)
            t = torch.randint(0, 1000, (B,), dtype=torch.int64)
            assert torch.equal(t, t_seen) and torch.equal(conds_seen, conds)
            pred = unet(noisy, t, conds).sample.detach()
            torch.manual_seed(seed)
            loss = me.training_step({"latents": latents, "conds": conds, "ids": list(range(B))}, 0)
            out[f"{ptype}_prior{prior}"] = {
                "prediction_type": ptype, "prior_loss_weight": prior, "seed": seed,
                "latents": latents, "conds": conds, "noise": noise, "timesteps": t, "noisy": noisy, "pred": pred,
                "loss_elem": loss_elem.detach().clone(), "loss": loss.detach().clone(), "logged": me.logged[-1]["train_loss"],
            }
    # the guards (model.py:324,332,336) and the unknown-type branch (:313-314)
    errs = {}
    me = shim.bare_lightning_module(model, unet=ToyUNet(), scheduler=RefDDIMScheduler("v_prediction"),
                                    config=shim.AttrDict({"prior_preservation": {"enabled": False}}))
    for name, batch in (("unknown_type", {"latents": torch.zeros(2, 4, 2, 2), "conds": torch.zeros(2, 3, 12)}),
                        ("nan_latents", {"latents": torch.full((2, 4, 2, 2), float("nan")), "conds": torch.zeros(2, 3, 12)}),
                        ("nan_conds", {"latents": torch.zeros(2, 4, 2, 2), "conds": torch.full((2, 3, 12), float("nan"))})):
        try:
            me.training_step(batch, 0)
            errs[name] = None
        except Exception as e:  # noqa: BLE001
            errs[name] = str(e)
    out["errors"] = errs
    torch.save(out, GOLDEN / "denoise_steps.pt")


def make_config_module_fixture():
    from types import SimpleNamespace

    from scal_sdt_b200.config import load_yaml
    from scal_sdt_b200.unet import UNet2DConditionModel, UNetConfig
    model = shim.load_reference_model()
    out = {"targets": {}}
    tdir = shim.REFERENCE_ROOT / "configs" / "optim_targets"
    for path in sorted(tdir.glob("*.yaml")):
        cfg = load_yaml(path)
        if cfg.get("unet") is None:
            continue
        torch.manual_seed(0)
        unet = UNet2DConditionModel(UNetConfig.tiny())
        groups = model.config_module(unet, cfg["unet"]["targets"])
        names = {id(p): n for n, p in unet.named_parameters()}
        rec = [{"params": [names[id(p)] for p in g["params"]], "overrides": {k: v for k, v in g.items() if k != "params"}}
               for g in groups]
        # on_save_checkpoint (model.py:378-391) on a module that holds this unet
        me = shim.bare_lightning_module(model, unet=unet)
        me.unet_ema = None
        ckpt = {}
        me.on_save_checkpoint(ckpt)
        out["targets"][path.stem] = {
            "groups": rec, "trainable": sorted(n for n, p in unet.named_parameters() if p.requires_grad),
            "injected": sorted(n for n, m in unet.named_modules() if hasattr(m, "lora_A")),
            "checkpoint_keys": sorted(ckpt["state_dict"].keys()),
        }
    # get_optimizer (model.py:33-64): AdamW by class name, beta1/beta2 -> betas, LR / weight-decay scaling
    scal = {}
    for method in ("sqrt", "linear"):
        p1, p2 = nn.Parameter(torch.zeros(3)), nn.Parameter(torch.zeros(2))
        conf = shim.AttrDict({"batch_size": 4, "optimizer": {
            "name": "torch.optim.AdamW", "params": {"lr": 5e-4, "beta1": 0.9, "beta2": 0.999, "weight_decay": 2e-2, "eps": 1e-7},
            "lr_scale": {"enabled": True, "method": method}}})
        trainer = SimpleNamespace(accumulate_grad_batches=2, num_nodes=1, num_devices=8)
        opt = model.get_optimizer([{"params": [p1], "lr": 1e-4, "weight_decay": 1e-2}, {"params": [p2]}], conf, trainer)
        scal[method] = [{"lr": g["lr"], "weight_decay": g["weight_decay"], "betas": list(g["betas"]), "eps": g["eps"]}
                        for g in opt.param_groups]
        scal[method + "_class"] = type(opt).__name__
    out["get_optimizer"] = {"accumulate": 2, "batch_size": 4, "nodes": 1, "devices": 8, "result": scal}
    (GOLDEN / "config_module.json").write_text(json.dumps(out))


def main():
    if not shim.available():
        raise SystemExit("/root/reference is not present: golden fixtures can only be generated in the build container")
    GOLDEN.mkdir(parents=True, exist_ok=True)
    make_bucket_fixtures()
    make_sampler_fixture()
    make_walker_fixture()
    make_ema_fixture()
    make_lora_glue_fixture()
    make_denoise_fixture()
    make_config_module_fixture()
    for f in sorted(GOLDEN.iterdir()):
        print(f"{f.name}: {f.stat().st_size} bytes")


if __name__ == "__main__":
    main()
