#!/usr/bin/env python
"""Benchmark of the LoRA train-step hot path (BASELINE.json metric: LoRA train latents/sec, SD1.5, 512^2).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, one process per GPU)
    python bench.py --impl reference --steps K --warmup W    # reference arm: the CPU path on the host cores

Workload at every N (weak scaling, per-GPU work fixed): BASELINE configs[1] -- SD1.5-shaped UNet (859.5 M frozen
parameters, random init), LoRA rank 16 on all 12 targets of the stock ``lora`` optim_target (192 sites), per-GPU
batch 8 of 4x64x64 latents + 77x768 text embeddings, bf16 compute, fp32 LoRA masters, AdamW.  One step =
noise/target -> UNet forward -> MSE loss -> backward -> gradient all-reduce -> AdamW -> operand repack.

One JSON line on stdout (rank 0).  ``value`` = whole-job latents/s with inputs resident in HBM; ``e2e`` = the same
through the public API with pinned-host inputs copied in and the loss read back every step; ``roofline`` = the fused
LoRA forward GEMM launches of the timed workload measured with CUDA events inside this process; ``cpu_baseline`` =
the oracle port of the reference's torch path on the host cores (bounded sample: batch 1 of the same workload).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

if "reference" in sys.argv:
    # torchrun exports OMP_NUM_THREADS=1 for every rank; the reference arm (rank 0 only) is the CPU path on ALL host
    # cores, so the thread-count variables must be fixed before torch (OpenMP / MKL) initialises
    _n = os.cpu_count() or 1
    try:
        import psutil
        _n = psutil.cpu_count(logical=False) or _n
    except Exception:  # noqa: BLE001
        pass
    for _k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = str(_n)

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 114514  # the reference's own seed (configs/lora.yaml:11)
WORKLOAD = dict(name="cfg2: SD1.5 LoRA r16 (192 sites: attn q/k/v/out + FF + proj_in/out), batch 8/GPU, 64x64 latents, bf16",
                rank=16, alpha=1, batch=8, h=64, w=64, ctx_len=77, ctx_dim=768)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------------------
# clocks during the timed region (pynvml polling thread)
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int, period_s: float = 0.05):
        self.index, self.period = index, period_s
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:  # noqa: BLE001
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)

    def summary(self):
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------------
# synthetic data
# ------------------------------------------------------------------------------------------------------------
def synthetic_batches(n: int, batch: int, rank: int, pin: bool):
    g = torch.Generator().manual_seed(SEED + 1000 * rank)
    out = []
    for _ in range(n):
        lat = torch.randn(batch, 4, WORKLOAD["h"], WORKLOAD["w"], generator=g)
        cond = torch.randn(batch, WORKLOAD["ctx_len"], WORKLOAD["ctx_dim"], generator=g)
        if pin:
            lat, cond = lat.pin_memory(), cond.pin_memory()
        out.append({"latents": lat, "conds": cond})
    return out


# ------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port of the reference's torch path, on the host cores
# ------------------------------------------------------------------------------------------------------------
def run_cpu_reference(steps: int, warmup: int, sample_batch: int = 1):
    from oracle.ref_trainer import RefTrainer
    from scal_sdt_b200.targets import lora_unet_targets
    from scal_sdt_b200.unet import UNet2DConditionModel, UNetConfig
    cores = os.cpu_count() or 1
    try:
        import psutil
        cores = psutil.cpu_count(logical=False) or cores        # the reference's own core metric (utils/sysinfo.py:4-5)
    except Exception:  # noqa: BLE001
        pass
    torch.set_num_threads(cores)
    torch.manual_seed(SEED)
    unet = UNet2DConditionModel(UNetConfig.sd15())
    tr = RefTrainer(unet, lora_unet_targets(WORKLOAD["rank"], WORKLOAD["alpha"]))
    g = torch.Generator().manual_seed(SEED)
    batches = synthetic_batches(2, sample_batch, 0, pin=False)

    def one(i):
        b = batches[i % len(batches)]
        noise = torch.randn(b["latents"].shape, generator=g)
        t = torch.randint(0, 1000, (sample_batch,), generator=g, dtype=torch.int64)
        return tr.step(b, noise, t)

    for i in range(warmup):
        one(i)
    t0 = time.perf_counter()
    for i in range(steps):
        loss = one(i)
    dt = time.perf_counter() - t0
    return dict(value=sample_batch * steps / dt, seconds=dt, cores=torch.get_num_threads(), loss=float(loss),
                sample=f"{steps} step(s) of batch {sample_batch} x 4x64x64 latents, same UNet/LoRA config, fp32 torch CPU "
                       f"(oracle port: loralib/diffusers/lightning are not installable offline)")


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    r = run_cpu_reference(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "LoRA train latents/sec SD1.5 512^2", "value": r["value"], "unit": "latents/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD["name"] + " [reference arm: CPU, bounded sample batch 1]"},
        "cpu_baseline": {"value": r["value"], "unit": "latents/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "latents/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def build_trainer(device, exchange):
    from scal_sdt_b200 import NoiseScheduler
    from scal_sdt_b200.targets import lora_unet_targets
    from scal_sdt_b200.trainer import LatentDiffusionTrainer
    from scal_sdt_b200.unet import UNet2DConditionModel, UNetConfig
    torch.manual_seed(SEED)                                   # identical init on every rank (DDP broadcast equivalent)
    with torch.device(device):
        unet = UNet2DConditionModel(UNetConfig.sd15())
    unet = unet.to(torch.bfloat16)                                          # frozen base held in bf16
    if os.environ.get("SDT_UNET_FORMAT", "nhwc") == "nhwc":
        unet = unet.to(memory_format=torch.channels_last)
    else:
        unet.channels_last = False
    sched = NoiseScheduler(prediction_type="epsilon")
    tr = LatentDiffusionTrainer(unet, sched, lora_unet_targets(WORKLOAD["rank"], WORKLOAD["alpha"]),
                                optimizer_params={"lr": 5e-4, "beta1": 0.9, "beta2": 0.999, "weight_decay": 2e-2, "eps": 1e-7},
                                lr_scale={"enabled": True, "method": "sqrt"}, batch_size=WORKLOAD["batch"],
                                exchange=exchange, seed=SEED)
    return tr


def site_flops(records):
    # one record per lora_gemm* launch; G = same-shape projections computed by that launch (grouped q/k/v, k/v)
    fwd = sum(G * (2.0 * M * K * N + 2.0 * M * R * (K + N)) for kind, M, K, N, R, G, *_ in records if kind == "fwd")
    bwd = sum(G * ((2.0 * M * K * N if dx else 0.0) + 4.0 * M * R * (K + N)) for kind, M, K, N, R, G, dx, *_ in records if kind == "bwd")
    return fwd, bwd


def elementwise_roofline(device, peaks):
    """HBM-bound kernels on arenas larger than L2 (in situ they are launch-latency bound: 131k elements)."""
    from scal_sdt_b200 import _lib
    lib = _lib.load()
    out = {}
    n = 256 * 1024 * 1024
    s = torch.empty(n, dtype=torch.float32, device=device).normal_()
    p = torch.empty(n, dtype=torch.float32, device=device).normal_()

    def timed(fn, iters=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e-3

    st = torch.cuda.current_stream().cuda_stream
    t = timed(lambda: _lib.check(lib.sdt_ema_update_flat(s.data_ptr(), p.data_ptr(), n, 0.005, None, 0, st)))
    out["ema_update_flat"] = {"gbs": 12.0 * n / t / 1e9, "frac": 12.0 * n / t / 1e9 / peaks["hbm"], "bytes": 12 * n}
    del s, p
    B, chw = 4096, 4 * 64 * 64          # 64 Mi elements per tensor
    x0 = torch.randn(B, chw, device=device)
    eps = torch.randn(B, chw, device=device)
    tt = torch.randint(0, 1000, (B,), device=device)
    noisy, v = torch.empty_like(x0), torch.empty_like(x0)
    from scal_sdt_b200 import scaled_linear_alphas_cumprod
    ac = scaled_linear_alphas_cumprod().to(device)
    t = timed(lambda: _lib.check(lib.sdt_noise_target(x0.data_ptr(), eps.data_ptr(), tt.data_ptr(), ac.data_ptr(), 1000,
                                                      noisy.data_ptr(), v.data_ptr(), 2, B, chw, 0, None, st)))
    nb = 4.0 * 4 * B * chw
    out["noise_target_v"] = {"gbs": nb / t / 1e9, "frac": nb / t / 1e9 / peaks["hbm"], "bytes": nb}
    ws = torch.zeros(lib.sdt_mse_loss_workspace_bytes(), dtype=torch.uint8, device=device)
    lo = torch.empty(3, device=device)
    t = timed(lambda: _lib.check(lib.sdt_mse_loss(noisy.data_ptr(), 0, v.data_ptr(), 0, lo.data_ptr(), x0.data_ptr(), None,
                                                  None, B, chw, B, 0.0, 1.0, ws.data_ptr(), st)))
    nb = 4.0 * 3 * B * chw
    out["mse_loss_dpred"] = {"gbs": nb / t / 1e9, "frac": nb / t / 1e9 / peaks["hbm"], "bytes": nb}
    return out


def ours_main(args):
    # Native libraries (NCCL's version banner, cuDNN warnings) write to fd 1; the contract is ONE JSON line on stdout,
    # so everything else is routed to stderr and the line is written to the saved descriptor at the end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch.distributed as dist
    from scal_sdt_b200 import GradExchange, _lib, build
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a CUDA device: the product path has no CPU fallback")
    if not build.LIB_PATH.exists():
        build.build_library()
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    exchange = GradExchange.from_torch_distributed(device)
    lib = _lib.load()
    peaks = load_peaks()
    B = WORKLOAD["batch"]

    tr = build_trainer(device, exchange)
    n_lora = sum(p.numel() for p, _, _ in tr.arena.slots)
    host = synthetic_batches(4, B, rank, pin=True)
    dev = [{k: v.to(device) for k, v in b.items()} for b in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up (eager), then capture the whole step as one CUDA graph ----
    for i in range(max(args.warmup, 3)):
        tr.step(dev[i % len(dev)])
    barrier()
    use_graph = not args.no_graph and not args.profile_step
    if use_graph:
        tr.enable_cuda_graph(dev[0])
        for i in range(2):
            tr.graphed_step(dev[i % len(dev)])
        barrier()
    do_step = tr.graphed_step if use_graph else tr.step

    if args.profile_step:
        torch.cuda.profiler.start()
        tr.step(dev[0])
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        os.write(real_stdout, (json.dumps({"profiled": "one training step", "workload": WORKLOAD["name"]}) + "\n").encode())
        exchange.close()
        return 0

    # ---- value: device-resident inputs, no host sync inside the region ----
    launches0 = lib.sdt_launch_count()
    with ClockSampler(local_rank) as clocks:
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            loss = do_step(dev[i % len(dev)])
        e1.record()
        barrier()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = lib.sdt_launch_count() - launches0        # eager launches (noise/timestep refresh are torch ops: not counted)
    if use_graph:
        launches += args.steps * tr.graph_launches_per_step     # our kernels inside the captured step, replayed K times
    value = world * B * args.steps / (ms_total * 1e-3)
    final_loss = float(loss.item())

    # ---- e2e: every step copies its inputs from pinned host memory and reads its loss back to the host ----
    # The read-back is a non-blocking D2H copy into a pinned slot per step; the region ends with a synchronize, so all
    # K results are on the host inside the timed region without draining the launch pipeline after every step.
    loss_host = torch.zeros(args.steps, dtype=torch.float32).pin_memory()
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        hb = host[i % len(host)]
        if use_graph:       # graphed_step copies the pinned host batch straight into the graph's static input buffers
            batch = hb
        else:
            batch = {k: v.to(device, non_blocking=True) for k, v in hb.items()}
        loss_host[i:i + 1].copy_(do_step(batch).detach().reshape(1), non_blocking=True)
    t1.record()
    barrier()
    ms_e2e = max_over_ranks(t0.elapsed_time(t1))
    step_loss = float(loss_host[-1])
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())

    # ---- roofline of the dominant kernel: the fused LoRA GEMM launches of a real training step ----
    # Pass 1 (eager, every rank: the all-reduce is collective): the ordered list of sites (kind, M, K, N, R) plus CUDA-event
    # brackets.  Once the step is faster than the host can launch it, event brackets also contain the host's launch gap,
    # so pass 2 takes the per-kernel durations of one replayed step from CUPTI (torch.profiler) and maps them onto the
    # site list by launch order (forward sites first, then backward in reverse order; the order is deterministic).
    from scal_sdt_b200 import lora as lora_mod
    roof = None
    lora_mod.PROFILE = [] if rank == 0 else None
    tr.step(dev[0])
    torch.cuda.synchronize()
    rec = lora_mod.PROFILE
    lora_mod.PROFILE = None
    kernel_times = None
    if rank == 0 or world > 1:
        try:
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                do_step(dev[1 % len(dev)])
                torch.cuda.synchronize()
            ks = []
            for e in prof.events():
                if "lora_gemm" in e.name and str(getattr(e, "device_type", "")).endswith("CUDA"):
                    ks.append((e.time_range.start, e.time_range.end - e.time_range.start))
            ks.sort()
            kernel_times = [d * 1e-6 for _, d in ks]      # seconds
        except Exception as exc:  # noqa: BLE001
            print(f"[bench] CUPTI pass unavailable ({exc}); falling back to CUDA-event brackets", file=sys.stderr)
    if rank == 0:
        fwd_sites = [r for r in rec if r[0] == "fwd"]
        bwd_sites = [r for r in rec if r[0] == "bwd"]
        n_fwd = len(fwd_sites)
        timing = "cuda_events"
        fwd_t = [a.elapsed_time(b) * 1e-3 for *_, a, b in fwd_sites]
        bwd_gemm_t = None
        t_bwd = sum(a.elapsed_time(b) for *_, a, b in bwd_sites) * 1e-3
        if kernel_times is not None and len(kernel_times) == n_fwd + len(bwd_sites):
            timing = "cupti_kernel_durations"
            fwd_t = kernel_times[:n_fwd]
            bwd_gemm_t = kernel_times[n_fwd:]
        f_fwd, f_bwd = site_flops(rec)
        t_fwd = sum(fwd_t)
        ach = f_fwd / t_fwd / 1e12
        # per-shape view of the same forward launches: which bound applies to which class of site
        shapes = {}
        for (kind, M, K, N, R, G, *_), sec in zip(fwd_sites, fwd_t):
            e = shapes.setdefault((M, K, N, R, G), [0, 0.0])
            e[0] += 1
            e[1] += sec
        by_shape = []
        for (M, K, N, R, G), (cnt, sec) in sorted(shapes.items(), key=lambda kv: -kv[1][1]):
            fl = G * (2.0 * M * K * N + 2.0 * M * R * (K + N))
            by = 2.0 * (M * K + G * (K * N + R * (K + N) + M * N + M * R))      # a grouped launch reads its shared X once
            tf, gbs = fl * cnt / sec / 1e12, by * cnt / sec / 1e9
            by_shape.append({"M": M, "K": K, "N": N, "R": R, "projections_per_launch": G, "launches": cnt, "avg_us": 1e6 * sec / cnt, "tflops": tf,
                             "frac_tensor": tf / peaks["tf_sustained"], "algorithmic_gbs": gbs, "frac_hbm": gbs / peaks["hbm"],
                             "bound": "tensor" if fl / by > peaks["tf_sustained"] * 1e3 / peaks["hbm"] else "hbm"})
        # DRAM bytes per GEMM launch from the committed ncu pass over one training step (dram__bytes_read + write summed over
        # the lora_gemm* launches / their count); bench.py cannot run ncu itself
        traffic, traffic_src = None, None
        for tname in ("r01_v21_kernels_per_step_ncu.json", "r01_v15_kernels_per_step_ncu.json", "r01_v5_lora_kernels_per_step_ncu.json"):
            tpath = os.path.join(ROOT, "profiles", tname)
            if not os.path.exists(tpath):
                continue
            with open(tpath) as f:
                tj = json.load(f)
            gl = [v for k, v in tj.items() if "lora_gemm" in k]
            if gl:
                traffic = sum((v["dram_read_MB"] + v["dram_write_MB"]) * 1e6 for v in gl) / sum(v["launches"] for v in gl)
                traffic_src = f"profiles/{tname} (avg over the forward + dX GEMM launches of a step)"
                break
        alg_bytes = sum(2.0 * (M * K + G * (K * N + R * (K + N) + M * N + M * R)) for kind, M, K, N, R, G, *_ in fwd_sites) / max(n_fwd, 1)
        dx_flops = sum(G * (2.0 * M * K * N + 2.0 * M * R * (K + N)) for kind, M, K, N, R, G, dx, *_ in bwd_sites if dx) + \
            sum(G * 2.0 * M * R * N for kind, M, K, N, R, G, dx, *_ in bwd_sites if not dx)
        backward = {"kernels": "lora_gemm* (dX, G) + lora_wgrad_kernel (dA and dB in one launch)", "timing": "cuda_events",
                    "achieved": f_bwd / t_bwd / 1e12, "frac": f_bwd / t_bwd / 1e12 / peaks["tf_sustained"]}
        if bwd_gemm_t is not None:
            backward["dx_gemm"] = {"timing": timing, "achieved": dx_flops / sum(bwd_gemm_t) / 1e12,
                                   "frac": dx_flops / sum(bwd_gemm_t) / 1e12 / peaks["tf_sustained"],
                                   "seconds_per_step": sum(bwd_gemm_t)}
        roof = {"kernel": "lora_gemm_kernel / lora_gemm_pair_kernel (K1: fused X W^T + bias + s (X A^T) B^T), every forward launch of a step (192 projections; q/k/v and cross-attention k/v go out as grouped launches)",
                "bound": "tensor", "achieved": ach, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                "frac": ach / peaks["tf_sustained"], "traffic": traffic, "traffic_source": traffic_src, "timing": timing,
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peaks["source"] + ", sustained bf16",
                "launches_timed": n_fwd, "avg_launch_us": 1e6 * t_fwd / max(n_fwd, 1),
                "flops_per_step": f_fwd, "forward_gemm_seconds_per_step": t_fwd,
                "backward": backward, "by_shape": by_shape, "elementwise": elementwise_roofline(device, peaks)}

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = run_cpu_reference(steps=2, warmup=1)
        cpu = {"value": r["value"], "unit": "latents/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    if rank == 0:
        line = {
            "metric": "LoRA train latents/sec SD1.5 512^2", "value": value, "unit": "latents/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD["name"], "global_batch": world * B, "lora_params": n_lora,
                       "parallelism": f"dp{world}", "optimizer": "fused AdamW over the flat LoRA arena",
                       "gradient_checkpointing": False, "cuda_graph": bool(use_graph),
                       "l2": "no flush: a step streams > 10 GB of weights/activations, far beyond the 126 MB L2"},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": "latents/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "final_loss": final_loss, "e2e_last_loss": step_loss,
            "roofline": roof, "cpu_baseline": cpu,
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    # teardown: the captured graph references the communicator, so it goes first
    tr.release_cuda_graph()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        # every rank has its result out; communicator teardown after graph capture has been seen to block, and there is
        # nothing left to flush but the process itself
        sys.stderr.flush()
        os._exit(0)
    exchange.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying one CUDA graph")
    ap.add_argument("--profile-step", action="store_true",
                    help="after warm-up run ONE step between cudaProfilerStart/Stop and exit (for ncu --profile-from-start off)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_main(args)
    return ours_main(args)


if __name__ == "__main__":
    sys.exit(main())
