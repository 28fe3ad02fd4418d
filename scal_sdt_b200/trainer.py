"""Training step of the hot path -- the shape of ``LatentDiffusionModel.training_step`` /
``_denoise_loss`` / ``on_train_batch_end`` (``modules/model.py:289-348,399-412``) without Lightning.

``LatentDiffusionTrainer`` owns: the UNet with injected LoRA modules, the flat parameter/gradient arena, the fused
AdamW(+EMA) step, the noising / loss kernels and the data-parallel gradient exchange.  One optimisation step is

    noise_and_target (1 launch) -> UNet forward (LoRA sites: 1 fused launch each) -> mse_loss (1 launch, emits dPred)
    -> backward (LoRA sites: dX + dA + dB launches) -> all-reduce of the gradient arena (1 NCCL call)
    -> AdamW + EMA over the arena (1 launch per hyper-parameter group) -> repack bf16 LoRA operands (1 launch)

and involves no host synchronisation: the loss value and the NaN flag stay on the device until asked for.
"""
from __future__ import annotations

import math
from typing import Any, Optional

import torch
from torch import nn

from .arena import LoraArena, ParamArena
from .comm import GradExchange
from .diffusion import DenoiseLoss, NoiseScheduler
from .ema import ExponentialMovingAverage
from .lora import deferred_wgrad, lora_modules
from .module_config import config_module, freeze_permanently
from .optim import FlatAdamW


def scale_lr(param_groups: list[dict], defaults: dict, batch_size: int, devices: int, nodes: int = 1, accumulate: int = 1,
             method: str = "sqrt") -> float:
    """``get_optimizer`` LR auto-scale (``modules/model.py:44-62``): lr *= c, weight_decay /= c,
    c = accumulate * batch * nodes * devices (sqrt of it for ``method == 'sqrt'``)."""
    coeff = accumulate * batch_size * nodes * devices
    if method == "sqrt":
        coeff = math.sqrt(coeff)
    elif method != "linear":
        raise ValueError(method)
    for g in param_groups:
        if "lr" in g:
            g["lr"] *= coeff
        if "weight_decay" in g:
            g["weight_decay"] /= coeff
    return coeff


class LatentDiffusionTrainer:
    def __init__(self, unet: nn.Module, scheduler: NoiseScheduler, unet_targets: Optional[list] = None, *,
                 optimizer_params: Optional[dict] = None, lr_scale: Optional[dict] = None, batch_size: int = 1,
                 prior_preservation: Optional[dict] = None, ema: Optional[dict] = None,
                 exchange: Optional[GradExchange] = None, seed: Optional[int] = None,
                 autocast_dtype: Optional[torch.dtype] = None):
        self.unet = unet
        # native fine-tune: fp32 master weights ARE the module's parameters and the host model runs under torch.autocast, the
        # reference's own mode (Lightning ``precision: 16 | bf16``, configs/__reserved_default__.yaml:44)
        self.autocast_dtype = autocast_dtype
        self.replayed_launches = 0           # libsdt_b200 kernels launched through graph replays so far
        self.scheduler = scheduler
        self.exchange = exchange or GradExchange(0, 1)
        self.device = next(unet.parameters()).device
        # ---- module injection / target selection (model.py:224-242) ----
        if unet_targets is not None:
            param_groups = config_module(unet, unet_targets)
        else:
            param_groups = [{"params": [p for p in unet.parameters() if p.requires_grad]}]
        if not any(len(g["params"]) for g in param_groups):
            freeze_permanently(unet)
            raise ValueError("no trainable parameters selected")
        self.param_groups_config = param_groups
        has_lora = any(True for _ in lora_modules(unet))
        self.arena: ParamArena = LoraArena(unet, param_groups) if has_lora else ParamArena(param_groups)
        # ---- optimizer (model.py:33-64) ----
        op = dict(optimizer_params or {"lr": 5e-4, "beta1": 0.9, "beta2": 0.999, "weight_decay": 2e-2, "eps": 1e-7})
        betas = (op.pop("beta1", 0.9), op.pop("beta2", 0.999))
        self.optimizer = FlatAdamW(self.arena, lr=op.get("lr", 1e-3), betas=op.get("betas", betas), eps=op.get("eps", 1e-8),
                                   weight_decay=op.get("weight_decay", 1e-2))
        if lr_scale and lr_scale.get("enabled", False):
            scale_lr(self.optimizer.param_groups, self.optimizer.defaults, batch_size, self.exchange.world,
                     method=lr_scale.get("method", "sqrt"))
        # ---- loss / EMA ----
        pp = prior_preservation or {}
        self.criterion = DenoiseLoss(self.device, bool(pp.get("enabled", False)), float(pp.get("prior_loss_weight", 1.0)))
        self.unet_ema: Optional[ExponentialMovingAverage] = None
        if ema and ema.get("enabled", False):
            # the reference keeps EMA on global rank 0 only (model.py:399-401); parameters are identical on every rank
            # after the step, so computing it everywhere is equivalent and keeps ranks symmetric
            self.unet_ema = ExponentialMovingAverage(unet, float(ema.get("decay", 0.995)))
        self._has_dropout = any(m.lora_dropout_p > 0.0 for _, m in lora_modules(unet))
        self.generator = torch.Generator(device=self.device)
        if seed is not None:
            self.generator.manual_seed(seed + self.exchange.rank)
        self.global_step = 0

    # ---- modules/model.py:289-316 ----------------------------------------------------------------------
    def _new_dropout_masks(self) -> None:
        if self._has_dropout:
            from .lora import advance_dropout_seed
            advance_dropout_seed(self.device, self.generator)

    def _denoise_loss(self, latents, conds, noise=None, timesteps=None, want_elementwise=False):
        if noise is None:
            noise = torch.randn(latents.shape, dtype=latents.dtype, device=latents.device, generator=self.generator)
        if timesteps is None:
            timesteps = torch.randint(0, self.scheduler.config.num_train_timesteps, (latents.shape[0],), dtype=torch.int64,
                                      device=latents.device, generator=self.generator)
        noisy, target = self.scheduler.noise_and_target(latents, noise, timesteps)
        if self.autocast_dtype is not None:
            with torch.autocast("cuda", dtype=self.autocast_dtype):
                pred = self.unet(noisy, timesteps, conds).sample
        else:
            pred = self.unet(noisy, timesteps, conds).sample
        return self.criterion(pred.contiguous(), target, want_elementwise=want_elementwise)

    # ---- modules/model.py:318-348 ----------------------------------------------------------------------
    def training_step(self, batch: dict, batch_idx: int = 0, noise=None, timesteps=None) -> torch.Tensor:
        if "latents" not in batch or "conds" not in batch:
            raise NotImplementedError("this path consumes cached latents/conds (the VAE / text encoder are out of scope)")
        return self._denoise_loss(batch["latents"], batch["conds"], noise, timesteps)

    def enable_overlapped_exchange(self, chunk_bytes: int = 64 << 20) -> None:
        """Reduce the gradient arena in chunks on a side stream WHILE backward is still producing the earlier layers'
        gradients (``train.py:98-109``: what DDP's bucketed reducer does for the reference).  For a native fine-tune the
        payload is the whole UNet (3.46 GB fp32 at SD scale) and a serial all-reduce after backward is pure added
        step time.  Only for arenas whose gradients arrive through autograd (LoRA sites write theirs from the kernels; their
        27-96 MB go out as one all-reduce)."""
        from .comm import OverlappedExchange
        if self.exchange.world > 1 and not isinstance(self.arena, LoraArena):
            self._overlap = OverlappedExchange(self.arena, self.exchange, chunk_bytes)

    def _exchange_gradients(self) -> None:
        ov = getattr(self, "_overlap", None)
        if ov is not None:
            ov.finish()
        else:
            self.exchange.all_reduce_mean_(self.arena.grads)

    def _backward(self, loss: torch.Tensor) -> None:
        """``loss.backward()`` with the LoRA weight-gradient reductions batched: a transformer block's dA / dB go out as one
        launch instead of one per site (``lora.deferred_wgrad``)."""
        if isinstance(self.arena, LoraArena):
            with deferred_wgrad():
                loss.backward()
        else:
            loss.backward()

    def optimizer_step(self) -> None:
        self._exchange_gradients()
        if self.unet_ema is not None and self.unet_ema._shadow_flat is not None:
            ema = self.unet_ema
            if not ema._shadow_flat.is_cuda:        # a caller mirrored the reference's ``unet_ema.to("cpu")`` (model.py:412)
                ema.to(self.device)
            if ema.num_updates is not None:
                ema.num_updates += 1
            self.optimizer.step(ema_shadow=ema._shadow_flat, ema_one_minus_decay=ema.current_one_minus_decay())
        else:
            self.optimizer.step()
            if self.unet_ema is not None:
                self.unet_ema.update()
        if isinstance(self.arena, LoraArena):
            self.arena.pack()

    def step(self, batch: dict, noise=None, timesteps=None) -> torch.Tensor:
        """zero_grad -> training_step -> backward -> exchange -> optimizer (+EMA) -> repack.  Returns the loss tensor
        (device resident; call ``.item()`` only when logging)."""
        self.optimizer.zero_grad()
        self._new_dropout_masks()
        loss = self.training_step(batch, self.global_step, noise, timesteps)
        if getattr(self, "_overlap", None) is not None:
            self._overlap.begin()
        self._backward(loss)
        self.optimizer_step()
        self.global_step += 1
        # detached: a caller that keeps the returned loss must not keep the step's autograd graph alive with it (its
        # AccumulateGrad nodes carry the eager stream and would poison a later CUDA-graph capture of the same parameters)
        return loss.detach()

    # ---- whole-step CUDA graph -------------------------------------------------------------------------------
    def enable_cuda_graph(self, example_batch: dict, warmup: int = 3) -> None:
        """Capture zero_grad -> forward -> loss -> backward -> all-reduce -> AdamW(+EMA) -> repack as ONE CUDA graph.

        The step launches ~3,000 kernels; once the GPU work per step drops towards the host's launch rate the Python /
        driver overhead shows.  (Drop references to losses of earlier ``training_step`` calls first: a live autograd graph keeps
        the parameters' AccumulateGrad nodes of the eager stream alive, and torch refuses to capture a backward that must
        synchronise with an uncaptured stream -- cudaErrorStreamCaptureIsolation.)  Replays read the batch, the noise and the timesteps from static buffers (filled by a few
        eager launches before each replay) and the optimizer's step-dependent scalars from a device table."""
        dev = self.device
        self._g_lat = example_batch["latents"].to(dev).clone()
        self._g_cond = example_batch["conds"].to(dev).clone()
        self._g_noise = torch.empty_like(self._g_lat)
        self._g_t = torch.zeros(self._g_lat.shape[0], dtype=torch.int64, device=dev)
        self._g_omd = torch.zeros(1, dtype=torch.float32, device=dev)
        from ._lib import PinnedRing
        self._g_omd_host = PinnedRing(self._g_omd)
        fused_ema = self.unet_ema is not None and self.unet_ema._shadow_flat is not None
        if self.unet_ema is not None and not fused_ema:
            raise NotImplementedError("CUDA-graph stepping needs the EMA shadow on the flat arena")

        def body():
            self.optimizer.zero_grad()
            loss = self._denoise_loss(self._g_lat, self._g_cond, self._g_noise, self._g_t)
            if getattr(self, "_overlap", None) is not None:
                self._overlap.begin()
            self._backward(loss)
            self._exchange_gradients()
            self.optimizer.step(use_device_hyper=True, ema_shadow=self.unet_ema._shadow_flat if fused_ema else None,
                                ema_one_minus_decay_dev=self._g_omd if fused_ema else None)
            if isinstance(self.arena, LoraArena):
                self.arena.pack()
            return loss.detach()

        # Warm-up replays the step for real (allocator, cuDNN autotune, tensor-map caches), so everything it trains is put
        # back afterwards: enabling the graph must not move the parameters, the moments, the EMA or any counter.
        saved = self._snapshot_train_state()
        self._refresh_step_inputs()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from . import _lib
        before = _lib.load().sdt_launch_count()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._g_loss = body()
        # kernels of libsdt_b200 recorded into the graph = launched on every replay
        self.graph_launches_per_step = int(_lib.load().sdt_launch_count() - before)
        self._restore_train_state(saved)

    def _snapshot_train_state(self) -> dict:
        ema = self.unet_ema
        return {"params": self.arena.params.clone(), "exp_avg": self.optimizer.exp_avg.clone(),
                "exp_avg_sq": self.optimizer.exp_avg_sq.clone(), "step_count": self.optimizer.step_count,
                "shadow": None if ema is None or ema._shadow_flat is None else ema._shadow_flat.clone(),
                "num_updates": None if ema is None else ema.num_updates, "rng": self.generator.get_state(),
                "global_step": self.global_step}

    def _restore_train_state(self, saved: dict) -> None:
        torch.cuda.synchronize()
        with torch.no_grad():
            self.arena.params.copy_(saved["params"])
            self.optimizer.exp_avg.copy_(saved["exp_avg"])
            self.optimizer.exp_avg_sq.copy_(saved["exp_avg_sq"])
            if saved["shadow"] is not None:
                self.unet_ema._shadow_flat.copy_(saved["shadow"])
        self.optimizer.step_count = saved["step_count"]
        if self.unet_ema is not None:
            self.unet_ema.num_updates = saved["num_updates"]
        self.generator.set_state(saved["rng"])
        self.global_step = saved["global_step"]
        self.arena.grads.zero_()
        if isinstance(self.arena, LoraArena):
            self.arena.pack()

    # ---- one graph per bucket shape (mixed-resolution batches) -----------------------------------------------------
    def enable_bucketed_cuda_graphs(self, example_batches: list, warmup: int = 2) -> None:
        """Aspect-ratio-bucketed training (``modules/dataset/bucket.py:154-207``) hands every rank a different latent shape per
        step, so one captured step does not do.  Per distinct ``(latents.shape, conds.shape)``: ONE graph of zero_grad ->
        noise/target -> forward -> loss -> backward, all sharing one memory pool (they never run concurrently).  The exchange
        and the update are shape-independent and stay outside: all-reduce + AdamW(+EMA) + repack are four launches.  No
        collective is captured or executed while a shape is being prepared, so ranks may meet new shapes at different
        steps."""
        dev = self.device
        if not hasattr(self, "_shape_graphs"):
            self._shape_graphs = {}
            self._graph_pool = torch.cuda.graph_pool_handle()
            self._g_t_shared = {}
            if not hasattr(self, "_g_omd"):
                from ._lib import PinnedRing
                self._g_omd = torch.zeros(1, dtype=torch.float32, device=dev)
                self._g_omd_host = PinnedRing(self._g_omd)
        for ex in example_batches:
            key = (tuple(ex["latents"].shape), tuple(ex["conds"].shape))
            if key in self._shape_graphs:
                continue
            ent = {"lat": ex["latents"].to(dev).clone(), "cond": ex["conds"].to(dev).clone()}
            ent["noise"] = torch.empty_like(ent["lat"])
            ent["t"] = torch.zeros(ent["lat"].shape[0], dtype=torch.int64, device=dev)

            def body(e=ent):
                self.optimizer.zero_grad()
                loss = self._denoise_loss(e["lat"], e["cond"], e["noise"], e["t"])
                self._backward(loss)
                return loss.detach()

            saved = self._snapshot_train_state()
            ent["noise"].normal_(generator=self.generator)
            ent["t"].random_(0, self.scheduler.config.num_train_timesteps, generator=self.generator)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    body()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            from . import _lib
            before = _lib.load().sdt_launch_count()
            ent["graph"] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(ent["graph"], pool=self._graph_pool):
                ent["loss"] = body()
            ent["launches"] = int(_lib.load().sdt_launch_count() - before)
            self._restore_train_state(saved)
            self._shape_graphs[key] = ent

    def bucketed_graphed_step(self, batch: dict) -> torch.Tensor:
        """``step`` for a batch whose shape was prepared by ``enable_bucketed_cuda_graphs``: replay that shape's forward /
        backward graph, then exchange + update eagerly."""
        key = (tuple(batch["latents"].shape), tuple(batch["conds"].shape))
        ent = self._shape_graphs.get(key)
        if ent is None:
            raise KeyError(f"no captured graph for batch shape {key}: pass an example to enable_bucketed_cuda_graphs first")
        ent["lat"].copy_(batch["latents"], non_blocking=True)
        ent["cond"].copy_(batch["conds"], non_blocking=True)
        ent["noise"].normal_(generator=self.generator)
        ent["t"].random_(0, self.scheduler.config.num_train_timesteps, generator=self.generator)
        self._new_dropout_masks()
        ent["graph"].replay()
        self.replayed_launches += ent["launches"]
        self.optimizer_step()
        self.global_step += 1
        return ent["loss"]

    def release_cuda_graph(self) -> None:
        """Drop the captured graphs (they pin the NCCL communicator and a private memory pool)."""
        if getattr(self, "_graph", None) is not None or getattr(self, "_shape_graphs", None):
            torch.cuda.synchronize()
        if getattr(self, "_graph", None) is not None:
            self._graph.reset()
            self._graph = None
        for ent in getattr(self, "_shape_graphs", {}).values():
            ent["graph"].reset()
        self._shape_graphs = {}

    def _refresh_step_inputs(self) -> None:
        """Eager, tiny: new noise / timesteps (modules/model.py:294,297-298) and the step-dependent optimizer scalars."""
        self._g_noise.normal_(generator=self.generator)
        self._g_t.random_(0, self.scheduler.config.num_train_timesteps, generator=self.generator)
        self._new_dropout_masks()
        self.optimizer.step_count += 1
        self.optimizer.refresh_device_hyper(self.optimizer.step_count)
        if self.unet_ema is not None:
            ema = self.unet_ema
            if ema.num_updates is not None:
                ema.num_updates += 1
            omd = ema.current_one_minus_decay()
            self._g_omd_host.push(lambda host: host.fill_(omd))

    def graphed_step(self, batch: dict) -> torch.Tensor:
        """Same contract as ``step`` (returns the device-resident loss), replaying the captured graph."""
        self._g_lat.copy_(batch["latents"], non_blocking=True)
        self._g_cond.copy_(batch["conds"], non_blocking=True)
        slot = batch.get("_host_slot")
        if slot is not None:                     # pinned staging slot of a CachedBatchLoader: free to be rewritten after this
            ev = torch.cuda.Event()
            ev.record()
            slot["_event"] = ev
        self._refresh_step_inputs()
        self._graph.replay()
        self.replayed_launches += self.graph_launches_per_step
        self.global_step += 1
        return self._g_loss

    # ---- modules/model.py:378-397 ----------------------------------------------------------------------
    def checkpoint_state_dict(self) -> dict[str, Any]:
        """Trainable-only state dict with the reference's keys (``unet.<path>.lora_A`` ..., plus ``unet_ema``)."""
        sd: dict[str, Any] = {f"unet.{n}": p.detach() for n, p in self.unet.named_parameters() if p.requires_grad}
        if self.unet_ema is not None:
            sd["unet_ema"] = self.unet_ema.state_dict()
        return sd

    def load_checkpoint_state_dict(self, sd: dict) -> None:
        if self.unet_ema is not None and "unet_ema" in sd:
            self.unet_ema.load_state_dict(sd["unet_ema"])
        own = dict(self.unet.named_parameters())
        with torch.no_grad():
            for k, v in sd.items():
                if k.startswith("unet.") and k[5:] in own:
                    own[k[5:]].copy_(v)
        from .lora import refresh_packed_operands
        refresh_packed_operands(self.unet)
