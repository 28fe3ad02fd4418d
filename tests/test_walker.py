"""CPU tier: target-selection DSL (``apply_module_config`` / ``config_module``) against the visits recorded from the
reference's own walker over every ``configs/optim_targets/*.yaml`` (golden ``walker.json``), and the injection /
checkpoint-key contract on the UNet skeleton."""
import json
from pathlib import Path

import torch
from torch import nn

from oracle.make_golden import clip_text_skeleton
from scal_sdt_b200 import LoRAConv2d, LoRALinear, apply_module_config, config_module, merge_config
from scal_sdt_b200.unet import UNet2DConditionModel, UNetConfig

GOLDEN = Path(__file__).parent / "golden"


def test_walker_visits_match_reference_for_every_optim_target():
    gold = json.loads((GOLDEN / "walker.json").read_text())
    assert {"lora", "lora_no-te", "lora_custom_diffusion", "custom_diffusion", "full_unet"} <= set(gold)
    for name, entry in gold.items():
        for comp, visits in entry["visits"].items():
            module = UNet2DConditionModel(UNetConfig.tiny()) if comp == "unet" else clip_text_skeleton()
            got = []

            def fn(sub, conf, path, _g=got):
                _g.append([path, type(sub).__name__, {k: v for k, v in dict(conf).items() if k not in ("targets", "index")}])
            apply_module_config(module, entry["config"][comp]["targets"], fn)
            assert got == visits, (name, comp)


def test_lora_yaml_selects_192_sites_and_injects():
    gold = json.loads((GOLDEN / "walker.json").read_text())["lora"]
    unet = UNet2DConditionModel(UNetConfig.tiny())
    groups = config_module(unet, gold["config"]["unet"]["targets"])
    assert len(groups) == 192                      # 16 transformer blocks x 12 modules
    assert all(g["lr"] == 5e-4 and g["weight_decay"] == 2e-2 and len(g["params"]) == 2 for g in groups)
    n_lin = sum(isinstance(m, LoRALinear) for m in unet.modules())
    n_conv = sum(isinstance(m, LoRAConv2d) for m in unet.modules())
    assert (n_lin, n_conv) == (160, 32)
    trainable = {n for n, p in unet.named_parameters() if p.requires_grad}
    assert len(trainable) == 384 and all(n.endswith(("lora_A", "lora_B")) for n in trainable)
    assert "down_blocks.0.attentions.0.transformer_blocks.0.attn2.to_k.lora_A" in trainable
    assert "mid_block.attentions.0.proj_in.lora_B" in trainable
    m = unet.get_submodule("up_blocks.3.attentions.2.transformer_blocks.0.ff.net.0.proj")
    assert m.r == 16 and m.scaling == 1 / 16 and int(m.lora_alpha) == 1
    assert "lora_alpha" in dict(m.named_buffers()) and "lora_alpha" not in dict(m.named_parameters())


def test_recurse_conf_accumulates_across_siblings():
    """``module.py:35-39``: recurse_conf merges into the running config and leaks to later siblings."""
    net = nn.ModuleDict({"a": nn.Linear(2, 2), "b": nn.Linear(2, 2), "c": nn.Linear(2, 2)})
    seen = {}
    apply_module_config(net, [{"index": ["a"], "recurse_conf": {"x": 1, "o": {"lr": 1}}},
                              {"index": ["b"], "recurse_conf": {"o": {"wd": 2}}},
                              {"index": ["c"], "y": 5}], lambda m, c, p: seen.__setitem__(p, c))
    assert seen["a"]["x"] == 1 and seen["b"]["o"] == {"lr": 1, "wd": 2} and seen["c"]["x"] == 1 and seen["c"]["y"] == 5


def test_merge_config_semantics():
    assert merge_config({"a": {"b": 1, "c": [1, 2]}}, {"a": {"c": [3]}, "d": 4}) == {"a": {"b": 1, "c": [3]}, "d": 4}


def test_sd15_sites_have_survey_shapes():
    """(K, N) of the 192 sites on the real SD1.5 widths (meta tensors: no memory)."""
    gold = json.loads((GOLDEN / "walker.json").read_text())["lora"]
    with torch.device("meta"):
        unet = UNet2DConditionModel(UNetConfig.sd15())
    shapes = {}

    def fn(sub, conf, path):
        k = (sub.in_features, sub.out_features) if isinstance(sub, nn.Linear) else (sub.in_channels, sub.out_channels)
        shapes[k] = shapes.get(k, 0) + 1
    apply_module_config(unet, gold["config"]["unet"]["targets"], fn)
    assert shapes[(320, 320)] == 40 and shapes[(320, 2560)] == 5 and shapes[(1280, 320)] == 5
    assert shapes[(640, 640)] == 40 and shapes[(1280, 1280)] == 48 and shapes[(768, 320)] == 10
    assert shapes[(768, 640)] == 10 and shapes[(768, 1280)] == 12 and shapes[(1280, 10240)] == 6
    assert sum(shapes.values()) == 192
    r = 16
    assert sum(n * r * (k[0] + k[1]) for k, n in shapes.items()) == 6_782_976      # SURVEY 8(e) LoRA param count


def test_kohya_export_key_layout():
    """``ckpt_tool.py:185-222``: lora_A -> lora_down.weight, lora_B -> lora_up.weight, alpha int32 from the run config."""
    from scal_sdt_b200.export import to_kohya_state_dict
    gold = json.loads((GOLDEN / "walker.json").read_text())["lora"]
    unet = UNet2DConditionModel(UNetConfig.tiny())
    config_module(unet, gold["config"]["unet"]["targets"])
    state = {f"unet.{n}": p.detach() for n, p in unet.named_parameters() if p.requires_grad}
    state["condition_model.encoder.text_model.encoder.layers.0.self_attn.q_proj.lora_A"] = torch.zeros(16, 8)
    state["condition_model.encoder.text_model.encoder.layers.0.self_attn.q_proj.lora_B"] = torch.zeros(8, 16)
    state["unet_ema"] = {"decay": 0.9}
    out = to_kohya_state_dict(state, optim_target=gold["config"])
    k = "lora_unet_down_blocks_0_attentions_0_transformer_blocks_0_attn1_to_q"
    assert out[k + ".lora_down.weight"].shape == (16, 32) and out[k + ".lora_up.weight"].shape == (32, 16)
    assert out[k + ".lora_down.weight"].dtype == torch.float16
    assert out[k + ".alpha"].dtype == torch.int32 and int(out[k + ".alpha"]) == 1
    assert "lora_te_text_model_encoder_layers_0_self_attn_q_proj.lora_up.weight" in out
    assert len(out) == 3 * 192 + 3
    proj = "lora_unet_mid_block_attentions_0_proj_in.lora_down.weight"
    assert out[proj].dim() == 2                     # 1x1-conv LoRA factors are emitted 2-D


def test_svd_extraction_matches_reference_restatement():
    """``extract_lora.py:23-39,130-154``: factors, sqrt(rank/alpha) scaling, kohya keys, 2-D conv weights."""
    from math import sqrt

    from scal_sdt_b200.export import extract_lora_state_dict, lora_approx
    from scal_sdt_b200.targets import lora_unet_targets
    torch.manual_seed(0)
    base = UNet2DConditionModel(UNetConfig.tiny())
    tuned = UNet2DConditionModel(UNetConfig.tiny())
    tuned.load_state_dict(base.state_dict())
    # a known rank-3 update on one Linear and one 1x1 conv, noise elsewhere
    lin = tuned.down_blocks[0].attentions[0].transformer_blocks[0].attn1.to_q
    conv = tuned.down_blocks[0].attentions[0].proj_in
    u, v = torch.randn(lin.out_features, 3), torch.randn(3, lin.in_features)
    with torch.no_grad():
        lin.weight += u @ v
        conv.weight += (torch.randn(conv.out_channels, 2) @ torch.randn(2, conv.in_channels))[:, :, None, None]
    state = extract_lora_state_dict(tuned, base, lora_unet_targets(rank=4, alpha=2), dtype=torch.float32)
    # 192 sites at SD1.5 topology -> the tiny UNet has the same topology
    assert len(state) == 3 * 192
    key = "lora_unet_down_blocks_0_attentions_0_transformer_blocks_0_attn1_to_q"
    down, up, alpha = state[f"{key}.lora_down.weight"], state[f"{key}.lora_up.weight"], state[f"{key}.alpha"]
    assert down.shape == (4, lin.in_features) and up.shape == (lin.out_features, 4)
    assert alpha.dtype == torch.int32 and int(alpha) == 2
    # (alpha / rank) * up @ down reproduces the rank-3 difference
    delta = lin.weight.detach() - base.down_blocks[0].attentions[0].transformer_blocks[0].attn1.to_q.weight.detach()
    assert torch.allclose((2 / 4) * up @ down, delta, atol=1e-4)
    # literal restatement of the reference's lora_approx on the same delta
    ru, rs, rvt = torch.linalg.svd(delta)
    ref_up, ref_down = ru[:, :4] @ torch.diag(rs[:4]), rvt[:4, :]
    assert torch.allclose(ref_up @ ref_down, (up @ down) / (sqrt(4 / 2) ** 2), atol=1e-4)
    d2, u2 = lora_approx(delta, 4)
    assert torch.allclose(u2 @ d2, ref_up @ ref_down, atol=1e-4)
    ckey = "lora_unet_down_blocks_0_attentions_0_proj_in"
    assert state[f"{ckey}.lora_down.weight"].dim() == 2 and state[f"{ckey}.lora_up.weight"].shape == (conv.out_channels, 4)
    # untouched sites extract (numerically) nothing
    zkey = "lora_unet_mid_block_attentions_0_transformer_blocks_0_attn2_to_v"
    assert state[f"{zkey}.lora_up.weight"].abs().max() < 1e-6
