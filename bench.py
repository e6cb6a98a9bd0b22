#!/usr/bin/env python
"""bench.py — ArcFlow-FLUX 2-NFE 1024x1024 images/sec on N B200s (BASELINE.json metric).

A "step" is one pass of the hot path over one batch: the whole 2-NFE denoise loop (2 transformer
forwards + 2 analytic sampler steps) for 8 images per GPU at 1024x1024 with cached text embeds
(configs[1] of BASELINE.json), synthetic weights of the true FLUX.1-dev + ArcFlow-adapter shapes.

  python bench.py [--gpus N] [--steps K] [--warmup W]           our arm (one rank per GPU under torchrun)
  python bench.py --impl reference [...]                        the reference's CPU torch path (oracle port)
  python bench.py --impl torch_cuda [...]                       the same torch restatement on the B200 (cuBLASLt GEMMs +
                                                                SDPA): what the reference dispatches to on a GPU — the
                                                                kernels to beat
  python bench.py --mode train [...]                            BASELINE.json configs[3]: FLUX distillation iteration
                                                                (2 student + 8 teacher forwards, adapter backward, DDP
                                                                all-reduce, clip + AdamW + EMA), samples/s, weak scaling
  python bench.py --model qwen [...]                            BASELINE.json configs[2]

Prints ONE JSON line on rank 0. `value` = device-timed whole-job images/s with inputs resident in HBM;
`e2e` = the same metric through the public pipeline call with HOST (pinned) buffers, H2D and D2H inside
the timed region; `roofline` = the dominant kernel (tcgen05 GEMM) measured live with CUDA events;
`cpu_baseline` = the oracle port timed on the host cores on a bounded sample (rank 0, N=1 only).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "images_per_sec_1024px_2nfe"
UNIT = "images/s"


_RESULT_FD = None


def _claim_stdout():
    """stdout must carry exactly ONE JSON line, but libraries write there too (NCCL prints its version banner through C
    stdio). From here on file descriptor 1 points at stderr for everything — Python and C alike, including buffers flushed
    at exit — and the result line is written to the original stdout by `_emit`."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def _ncu_traffic():
    """DRAM bytes per launch of the two tensor-core kernels from the committed `ncu --set full` capture (dram__bytes_read.sum
    + dram__bytes_write.sum; profiles/r01_ncu_traffic.json names the launches) — a profiler figure cannot be taken live."""
    for name in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", name)
        try:
            with open(path) as f:
                d = json.load(f)
            d.setdefault("file", "profiles/" + name)
            return d
        except (OSError, ValueError):
            continue
    return {}


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(bf16=d.get("bf16_tflops_sustained", 1386.4), bf16_burst=d.get("bf16_tflops", 1653.1),
                    hbm=d.get("hbm_gbs", 6548.8), source="measured")
    return dict(bf16=1400.0, bf16_burst=1590.0, hbm=6650.0, source="fallback")


def flux_flops_per_image_nfe(S_img: int, S_txt: int = 512, lora_rank: int = 256) -> dict:
    """Algorithmic FLOPs per image per network call (SURVEY.md §8d / BASELINE.md §3)."""
    D, M, S = 3072, 12288, S_img + S_txt
    linear = 57 * 2 * S * (4 * D * D + 2 * D * M)
    attn = 57 * 4 * S * S * D
    r = lora_rank
    lora = S * (19 * 4 * r * (D + M) + 38 * (2 * r * (D + M) + 2 * r * (2 * D + M))) if r else 0
    heads = 2 * S_img * D * 1148
    embed = 2 * S_img * 64 * D + 2 * S_txt * 4096 * D
    return dict(linear=linear, attn=attn, lora=lora, heads=heads, embed=embed,
                total=linear + attn + lora + heads + embed)


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# =================================================================================================
# reference arm: the reference's own CPU torch path (oracle port) on the host cores, bounded sample
# =================================================================================================
def cpu_sample_images_per_sec(px: int, nfe: int, threads: int, dbl: int = 1, sgl: int = 2,
                              full=(19, 38)) -> dict:
    """EXTRAPOLATED figure for the metric's own configuration: times `dbl` double + `sgl` single FLUX blocks (of 19 + 38)
    of the reference's bf16 torch path for ONE image at `px` on the host cores, one forward, and scales by depth and NFE
    (a full-depth 1024^2 image is ~160 TFLOP = minutes per image on CPU; the sample keeps the run bounded).
    Embedders/heads/sampler (<0.1 % of the FLOPs) are included once, unscaled."""
    import torch
    from oracle import arcflow_oracle as O
    from arcflow_b200.config import ArcFluxConfig
    from arcflow_b200.synthetic import make_flux_state_dict, make_flux_inputs
    torch.set_num_threads(threads)
    cfg = ArcFluxConfig(num_layers=dbl, num_single_layers=sgl)
    sd = make_flux_state_dict(cfg, seed=1234, device="cpu")
    x, txt, pooled = make_flux_inputs(cfg, 1, px, px, seed=42)
    grid = (px // 16, px // 16)
    args = (x.bfloat16(), txt, pooled, torch.tensor([1.0]), torch.tensor([3.5]), grid)
    with torch.no_grad():
        t0 = time.perf_counter()
        out = O.flux_forward(sd, cfg, *args, dtype=torch.bfloat16)
        dt = time.perf_counter() - t0
    blocks_full, blocks_s = sum(full), dbl + sgl
    sec_per_image = dt * blocks_full / blocks_s * nfe
    return dict(value=1.0 / sec_per_image, seconds_sample=dt, extrapolated=True,
                sample=f"EXTRAPOLATED: 1 image, {dbl} double + {sgl} single of {full[0]}+{full[1]} FLUX blocks at {px}px "
                       f"(S={512 + grid[0] * grid[1]}), bf16 torch CPU, one forward, scaled x{blocks_full / blocks_s:.1f} "
                       f"depth x{nfe} NFE")


def cpu_full_depth_256(threads: int, nfe: int = 2, repeats: int = 1) -> dict:
    """MEASURED (not extrapolated): BASELINE.json configs[0] / BASELINE.md §4 — ArcFlow-FLUX 256x256, batch 1, FULL depth
    (19 + 38 blocks), the real `nfe`-NFE loop (transformer + momentum integration) of the oracle port in bf16 on the
    host cores. Seeding 12 B parameters with the CPU generator takes minutes and 24 GB, so every double (single) block
    ALIASES the tensors of block 0 — the arithmetic, the per-block weight traffic (340 / 141 MB, beyond any L2) and the
    loop are the real ones; only the values repeat, which timing does not see."""
    import torch
    from oracle import arcflow_oracle as O
    from arcflow_b200.config import ArcFluxConfig, flux_dev
    from arcflow_b200.synthetic import make_flux_state_dict, make_flux_inputs
    torch.set_num_threads(threads)
    one = make_flux_state_dict(ArcFluxConfig(num_layers=1, num_single_layers=1), seed=1234, device="cpu")
    cfg = flux_dev()
    sd = dict(one)
    for k, v in one.items():
        if k.startswith("transformer_blocks.0."):
            for i in range(1, cfg.num_layers):
                sd[f"transformer_blocks.{i}." + k[len("transformer_blocks.0."):]] = v
        elif k.startswith("single_transformer_blocks.0."):
            for i in range(1, cfg.num_single_layers):
                sd[f"single_transformer_blocks.{i}." + k[len("single_transformer_blocks.0."):]] = v
    x, txt, pooled = make_flux_inputs(cfg, 1, 256, 256, seed=42)
    best = None
    with torch.no_grad():
        for _ in range(max(1, repeats)):
            t0 = time.perf_counter()
            out = O.flux_denoise(sd, cfg, x, txt, pooled, (16, 16), num_inference_steps=nfe, dtype=torch.bfloat16)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    assert bool(torch.isfinite(out).all())
    return dict(images_per_sec=1.0 / best, seconds_per_image=best, cores=threads, nfe=nfe, extrapolated=False,
                workload="ArcFlow-FLUX 256x256 batch 1, 19+38 blocks, real 2-NFE loop, bf16 torch CPU (BASELINE.json configs[0])",
                tflop_per_image=flux_flops_per_image_nfe(256)["total"] * nfe / 1e12)


def _cpu_model_name() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if args.model != "flux" or args.mode != "infer":
        _emit({"impl": "reference", "unavailable": "the CPU reference arm is implemented for the FLUX inference headline only"})
        return 0
    import torch
    threads = os.cpu_count() or 1
    for _ in range(args.warmup if args.warmup < 1 else 1):
        cpu_sample_images_per_sec(args.px, args.nfe, threads)
    vals = [cpu_sample_images_per_sec(args.px, args.nfe, threads) for _ in range(max(1, min(args.steps, 3)))]
    best = max(vals, key=lambda d: d["value"])
    full256 = None
    if not args.no_cpu_full:
        try:
            full256 = cpu_full_depth_256(threads, 2)
        except Exception as ex:   # the extrapolated figure still stands
            full256 = {"unavailable": f"{type(ex).__name__}: {ex}"}
    line = {
        "impl": "reference", "metric": METRIC, "value": best["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * args.batch / best["value"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": _config(args, 1),
        "cpu_baseline": {"value": best["value"], "unit": UNIT, "cores": threads, "kind": "port", "extrapolated": True,
                         "cpu_model": _cpu_model_name(), "sample": best["sample"],
                         "measured_256px_b1_full_depth": full256},
        "e2e": {"value": best["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference deps (diffusers/peft/mmcv) are not installable offline; this is the oracle port of its "
                "torch path (oracle/arcflow_oracle.py) on the host cores. `value` is EXTRAPOLATED from a depth sample of one "
                "1024^2 image (see cpu_baseline.sample); cpu_baseline.measured_256px_b1_full_depth is a real full-depth "
                "2-NFE loop at BASELINE.json configs[0]. The same-GPU baseline is `--impl torch_cuda`.",
    }
    _emit(line)
    return 0


# =================================================================================================
# torch-CUDA arm: the reference's torch path on the SAME B200 (cuBLASLt + SDPA) — the kernels to beat
# =================================================================================================
def torch_cuda_denoise_fn(args, dev, rank: int = 0):
    """Returns (fn, info): fn() runs one batch of the 2-NFE loop with the oracle port in bf16 on `dev` — nn.Linear ->
    cuBLASLt, F.scaled_dot_product_attention (cuDNN / flash backend), unfused LN / GELU / RoPE / LoRA branches and the
    ~25 elementwise sampler kernels per NFE, i.e. what lakonlab/models/architecture/arcflow/arcflux.py:180-230 dispatches
    to on a GPU. Same synthetic weights (seed 1234) and inputs (seed 42 + rank) as our arm."""
    import torch
    from oracle import arcflow_oracle as O
    qwen = args.model == "qwen"
    grid = (args.px // 16, args.px // 16)
    if qwen:
        from arcflow_b200.qwen import make_qwen_inputs, make_qwen_state_dict, qwen_image
        cfg = qwen_image()
        sd = make_qwen_state_dict(cfg, seed=1234, device=dev)
        x, txt = make_qwen_inputs(cfg, args.batch, args.px, args.px, 512, 42 + rank, dev)
        run = lambda: O.qwen_denoise(sd, cfg, x, txt, grid, num_inference_steps=args.nfe, dtype=torch.bfloat16)
    else:
        from arcflow_b200.config import flux_dev
        from arcflow_b200.synthetic import make_flux_inputs, make_flux_state_dict
        cfg = flux_dev()
        sd = make_flux_state_dict(cfg, seed=1234, device=dev)
        x, txt, pooled = make_flux_inputs(cfg, args.batch, args.px, args.px, seed=42 + rank, device=dev)
        run = lambda: O.flux_denoise(sd, cfg, x, txt, pooled, grid, num_inference_steps=args.nfe, dtype=torch.bfloat16)
    backends = []
    try:
        from torch.nn.attention import SDPBackend, sdpa_kernel
        order = [SDPBackend.CUDNN_ATTENTION, SDPBackend.FLASH_ATTENTION, SDPBackend.EFFICIENT_ATTENTION, SDPBackend.MATH]
        backends = [b.name for b in order]

        def fn():
            with torch.no_grad(), sdpa_kernel(order, set_priority=True):
                return run()
    except Exception:
        def fn():
            with torch.no_grad():
                return run()
    info = dict(kernels="torch.nn.functional.linear (cuBLASLt bf16) + scaled_dot_product_attention + eager elementwise",
                sdpa_backend_priority=backends, torch=torch.__version__)
    return fn, info


def _time_cuda(fn, steps, warmup, world, dev, after=None):
    import torch
    import torch.distributed as dist
    for _ in range(warmup):
        out = fn()
        if after:
            after(out)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = fn()
        if after:
            after(out)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_torch_cuda(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.mode != "infer":
        if rank == 0:
            _emit({"impl": "torch_cuda", "unavailable": "the torch-CUDA arm covers the inference metric (FLUX / Qwen) only"})
        return 0
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl torch_cuda: no CUDA device")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    fn, info = torch_cuda_denoise_fn(args, dev, rank)
    gathered = [None]

    def after(out):
        if world > 1:
            if gathered[0] is None:
                gathered[0] = torch.empty(world * out.shape[0], *out.shape[1:], device=dev)
            dist.all_gather_into_tensor(gathered[0], out.contiguous())

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_total = _time_cuda(fn, args.steps, max(args.warmup, 3), world, dev, after)
    clock_info = clocks.stop() if rank == 0 else {}
    if rank == 0:
        images = world * args.batch * args.steps
        value = images / (ms_total / 1000.0)
        grid = (args.px // 16, args.px // 16)
        fl = (qwen_flops_per_image_nfe if args.model == "qwen" else flux_flops_per_image_nfe)(grid[0] * grid[1], 512, 256)
        step_ms = ms_total / args.steps
        peaks = _peaks()
        _emit({"impl": "torch_cuda", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": _config(args, world), "clocks": clock_info,
               "step_tflops": fl["total"] * args.nfe * args.batch / (step_ms * 1e9),
               "step_frac_of_peak": fl["total"] * args.nfe * args.batch / (step_ms * 1e9) / peaks["bf16"],
               "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "gpu_launches": 0, "library": info,
               "note": "none of this repo's kernels run in this arm: it is the oracle port of the reference's torch path "
                       "executed by PyTorch's own CUDA libraries on the same GPU"})
    if world > 1:
        dist.destroy_process_group()
    return 0


def qwen_flops_per_image_nfe(S_img: int, S_txt: int = 512, r: int = 256) -> dict:
    """Algorithmic FLOPs of one ArcFlow-Qwen-Image forward: 60 double-stream blocks, LoRA on the MLPs
    (text MLP of the last block excluded), heads and embedders."""
    D, M, S = 3072, 12288, S_img + S_txt
    linear = 60 * 2 * S * (4 * D * D + 2 * D * M)
    attn = 60 * 4 * S * S * D
    lora = 60 * S_img * 4 * r * (D + M) + 59 * S_txt * 4 * r * (D + M)
    return {"total": linear + attn + lora + 2 * S_img * D * 1148 + 2 * S_img * 64 * D + 2 * S_txt * 3584 * D}


def _config(args, world):
    if getattr(args, "model", "flux") == "qwen":
        return {"workload": f"ArcFlow-Qwen-Image {args.nfe}-NFE {args.px}x{args.px} batch {args.batch}/GPU, cached "
                            f"Qwen2.5-VL embeds (BASELINE.json configs[2])",
                "global_batch": args.batch * world, "txt_len": 512, "img_tokens": (args.px // 16) ** 2,
                "nfe": args.nfe, "shift": 3.2, "timestep_ratio": 1.0,
                "parallelism": f"batch-parallel dp{world}, weights replicated, all-gather of final latents",
                "weights": "synthetic Qwen-Image shapes (60 blocks, D=3072) + rank-256 ArcFlow adapter, un-merged",
                "l2": "working set (41 GB weights + activations) >> 126 MB L2; no explicit flush"}
    return {"workload": f"ArcFlow-FLUX {args.nfe}-NFE {args.px}x{args.px} batch {args.batch}/GPU, cached T5/CLIP embeds "
                        f"(BASELINE.json configs[1])",
            "global_batch": args.batch * world, "txt_len": 512, "img_tokens": (args.px // 16) ** 2,
            "nfe": args.nfe, "shift": 3.2, "timestep_ratio": 1.0, "guidance": 3.5,
            "parallelism": f"batch-parallel dp{world}, weights replicated, all-gather of final latents",
            "weights": "synthetic FLUX.1-dev shapes (19+38 blocks, D=3072) + rank-256 ArcFlow adapter, un-merged",
            "l2": "working set (25 GB weights + 2.3 GB activations per step) >> 126 MB L2; no explicit flush"}


# =================================================================================================
# our arm
# =================================================================================================
def run_ours(args):
    import torch
    import torch.distributed as dist
    from arcflow_b200 import _lib, build
    from arcflow_b200.config import flux_dev
    from arcflow_b200.model import ArcFluxEngineModel
    from arcflow_b200.synthetic import make_flux_state_dict, make_flux_inputs
    from lakonlab.pipelines.arcflux_pipeline import ArcFluxPipeline

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:       # one builder: concurrent ranks must not link the same .so
        build.build()
    if world > 1:
        dist.barrier()
    if rank != 0:
        build.build()   # no-op: the fingerprint matches what rank 0 just verified / built
    lib = _lib.load()

    qwen = args.model == "qwen"
    grid = (args.px // 16, args.px // 16)
    if qwen:   # BASELINE.json configs[2]; the headline (and the default) is FLUX
        from arcflow_b200.qwen import ArcQwenEngineModel, make_qwen_inputs, make_qwen_state_dict, qwen_image
        from lakonlab.pipelines.arcqwen_pipeline import ArcQwenImagePipeline
        cfg = qwen_image()
        sd = make_qwen_state_dict(cfg, seed=1234, device=dev)
        model = ArcQwenEngineModel(sd, cfg, device=dev, consume_state_dict=True)
        del sd
        torch.cuda.empty_cache()
        pipe = ArcQwenImagePipeline(transformer=model)
        x, txt = make_qwen_inputs(cfg, args.batch, args.px, args.px, 512, 42 + rank, dev)
        pooled = None
    else:
        cfg = flux_dev()
        sd = make_flux_state_dict(cfg, seed=1234, device=dev)
        model = ArcFluxEngineModel(sd, cfg, device=dev, consume_state_dict=True)
        del sd
        torch.cuda.empty_cache()
        pipe = ArcFluxPipeline(transformer=model)
        x, txt, pooled = make_flux_inputs(cfg, args.batch, args.px, args.px, seed=42 + rank, device=dev)
    gathered = torch.empty(world * args.batch, *x.shape[1:], device=dev) if world > 1 else None

    def denoise_once():
        if qwen:
            return model.denoise(x, txt, grid, num_inference_steps=args.nfe)
        return model.denoise(x, txt, pooled, grid, num_inference_steps=args.nfe)

    def step_device():
        out = denoise_once()
        if world > 1:   # sampler-boundary exchange: final packed latents of every rank
            dist.all_gather_into_tensor(gathered, out)
        return out

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    sync_all()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = lib.afb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = lib.afb_launch_count() - launches0
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    images = world * args.batch * args.steps
    value = images / (ms_total / 1000.0)

    # ---- e2e: the public pipeline call with pinned HOST buffers, copies inside the timed region ----
    # N > 1: every rank makes the SAME call with the GLOBAL batch; pipe(...) keeps this rank's images (sliced on the host
    # side of the H2D copy), denoises them and all-gathers the final latents (lakonlab/pipelines/arcflux_pipeline.py).
    if world > 1:
        gb = world * args.batch
        if qwen:
            gx, gtxt = make_qwen_inputs(cfg, gb, args.px, args.px, 512, 42, "cpu")
            gpooled = None
        else:
            gx, gtxt, gpooled = make_flux_inputs(cfg, gb, args.px, args.px, seed=42, device="cpu")
        host_in = [t_.pin_memory() for t_ in (gx, gtxt, gpooled) if t_ is not None]
    else:
        host_in = [t_.cpu().pin_memory() for t_ in (x, txt, pooled) if t_ is not None]
    hx, htxt = host_in[0], host_in[1]
    hpooled = host_in[2] if len(host_in) > 2 else None
    hout = torch.empty_like(hx).pin_memory()

    def step_e2e():
        if qwen:
            r = pipe(prompt_embeds=htxt, latents=hx, height=args.px, width=args.px, num_inference_steps=args.nfe,
                     timestep_ratio=1.0, output_type="latent")
        else:
            r = pipe(prompt_embeds=htxt, pooled_prompt_embeds=hpooled, latents=hx, height=args.px, width=args.px,
                     num_inference_steps=args.nfe, timestep_ratio=1.0, guidance_scale=3.5, output_type="latent")
        hout.copy_(r.images, non_blocking=True)

    step_e2e()
    sync_all()
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    sync_all()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = images / (float(t.item()) / 1000.0)
    clock_info = clocks.stop() if rank == 0 else {}
    # whole-job bytes per step: every rank copies in its own shard and reads back the gathered result
    h2d = sum(t_.numel() * t_.element_size() for t_ in host_in)
    d2h = world * hout.numel() * hout.element_size()

    # ---- roofline of the dominant kernel: one extra instrumented step (events around every launch) ----
    model.set_profiling(True)
    model.read_profile()
    denoise_once()
    prof = model.read_profile()
    model.set_profiling(False)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    peaks = _peaks()
    traffic = _ncu_traffic()
    fl = (qwen_flops_per_image_nfe if qwen else flux_flops_per_image_nfe)(grid[0] * grid[1], 512, cfg.lora_rank)
    step_flops = fl["total"] * args.nfe * args.batch
    gemm_tf = prof["gemm_flops"] / (prof["gemm_ms"] * 1e9) if prof["gemm_ms"] > 0 else 0.0
    attn_tf = prof["attn_flops"] / (prof["attn_ms"] * 1e9) if prof["attn_ms"] > 0 else 0.0
    step_ms = ms_total / args.steps
    # the kernel's launches by epilogue: the QKV launches whose epilogue also does the per-head RMSNorm + RoPE carry the work
    # of the former rmsnorm_rope kernel (114 launches x 0.19 ms per forward), so their time per FLOP is not a pure-GEMM figure
    fq_ms, fq_fl, fq_n = prof.get("gemm_fused_qk_ms", 0.0), prof.get("gemm_fused_qk_flops", 0.0), prof.get("gemm_fused_qk_launches", 0)
    pl_ms, pl_fl, pl_n = prof["gemm_ms"] - fq_ms, prof["gemm_flops"] - fq_fl, prof["gemm_launches"] - fq_n
    by_epilogue = {"plain_bias_gelu_gate_res": {"launches": pl_n, "ms": pl_ms, "achieved": pl_fl / (pl_ms * 1e9) if pl_ms > 0 else None,
                                                "frac": pl_fl / (pl_ms * 1e9) / peaks["bf16"] if pl_ms > 0 else None}}
    if fq_n:
        by_epilogue["qkv_with_fused_rmsnorm_rope"] = {
            "launches": fq_n, "ms": fq_ms, "achieved": fq_fl / (fq_ms * 1e9), "frac": fq_fl / (fq_ms * 1e9) / peaks["bf16"],
            "note": "GEMM FLOPs only over a launch that also normalises and rotates q / k in its epilogue (no rmsnorm_rope launches remain in the step)"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic", "config": _config(args, world),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "clocks": clock_info,
        "roofline": {
            "kernel": "gemm_bf16_2cta_kernel (tcgen05 cta_group::2, all Linear/LoRA launches of one step)",
            "bound": "tensor", "achieved": gemm_tf, "peak": peaks["bf16"], "unit": "TFLOP/s",
            "frac": gemm_tf / peaks["bf16"], "traffic": traffic.get("gemm", {}).get("dram_bytes_per_launch"),
            "traffic_launch": traffic.get("gemm", {}).get("launch"),
            "traffic_algorithmic_bytes": traffic.get("gemm", {}).get("algorithmic_bytes_per_launch"),
            "traffic_source": f"{traffic.get('file')}: {traffic.get('source')}" if traffic else None,
            "peak_source": f"{peaks['source']} bf16_tflops_sustained (kernel timed inside a long step)",
            "launches": prof["gemm_launches"], "avg_launch_ms": prof["gemm_ms"] / max(prof["gemm_launches"], 1),
            "algorithmic_flops_per_step": prof["gemm_flops"],
            "share_of_step": prof["gemm_ms"] / step_ms if step_ms else None,
            "by_epilogue": by_epilogue,
        },
        "roofline_attention": {
            "kernel": "attention_fwd_split_kernel (tcgen05; bounded-score softmax)", "bound": "tensor", "achieved": attn_tf,
            "peak": peaks["bf16"], "unit": "TFLOP/s", "frac": attn_tf / peaks["bf16"],
            "launches": prof["attn_launches"], "avg_launch_ms": prof["attn_ms"] / max(prof["attn_launches"], 1),
            "algorithmic_bytes_per_launch": 4 * (512 + grid[0] * grid[1]) * 3072 * 2 * args.batch,
            "traffic": (traffic.get("attention") or {}).get("dram_bytes_per_launch"),
            "share_of_step": prof["attn_ms"] / step_ms if step_ms else None,
        },
        "step_tflops": step_flops / (step_ms * 1e9), "step_frac_of_peak": step_flops / (step_ms * 1e9) / peaks["bf16"],
        "tflop_per_image": fl["total"] * args.nfe / 1e12,
    }
    if world == 1 and args.variants and not qwen:
        # Not the headline (BASELINE.json's metric is latents out): the public call with the native VAE decoder attached
        # and output_type='pt' — host embeds / latents in, decoded fp32 images on the host out, everything in the timed region.
        try:
            from arcflow_b200.vae import FluxVAEDecoder
            from arcflow_b200.synthetic import make_vae_decoder_state_dict  # seeded synthetic decoder weights (no checkpoint offline)
            pipe.vae = FluxVAEDecoder(make_vae_decoder_state_dict(seed=7, device=dev), device=dev)
            himg = torch.empty((args.batch, 3, args.px, args.px), dtype=torch.float32).pin_memory()

            def step_decode():
                r = pipe(prompt_embeds=htxt, pooled_prompt_embeds=hpooled, latents=hx, height=args.px, width=args.px,
                         num_inference_steps=args.nfe, timestep_ratio=1.0, guidance_scale=3.5, output_type="pt")
                himg.copy_(r.images, non_blocking=True)

            step_decode()
            torch.cuda.synchronize()
            v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            v0.record()
            for _ in range(args.steps):
                step_decode()
            v1.record()
            torch.cuda.synchronize()
            vms = v0.elapsed_time(v1) / args.steps
            line.setdefault("variants", {})["e2e_with_vae_decode"] = {
                "value": args.batch / (vms / 1000.0), "unit": UNIT, "ms_per_step": vms,
                "d2h_bytes_per_step": himg.numel() * 4,
                "vae_tflop_per_image": FluxVAEDecoder.flops(args.px // 8, args.px // 8) / 1e12,
                "note": "pipe(..., output_type='pt'): 2-NFE denoise + native FLUX VAE decode (implicit-GEMM convolutions on the "
                        "tcgen05 kernel) + fp32 image read-back; synthetic decoder weights"}
            pipe.vae = None
            del himg
        except Exception as ex:
            line.setdefault("variants", {})["e2e_with_vae_decode"] = {"unavailable": f"{type(ex).__name__}: {ex}"}
        torch.cuda.empty_cache()
    if world == 1 and args.variants:
        # Not the headline: the same step with the adapter merged into the base weights (pipe.fuse_lora(), one-way, so it
        # runs last). The LoRA branches are 5.6 % of the algorithmic FLOPs counted above; images/s is what a deployment
        # that accepts bf16-rounded merged weights gets.
        pipe.fuse_lora()
        for _ in range(2):
            step_device()
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            step_device()
        f1.record()
        torch.cuda.synchronize()
        line.setdefault("variants", {})["fuse_lora"] = {
            "value": args.batch * args.steps / (f0.elapsed_time(f1) / 1000.0), "unit": UNIT,
            "note": "adapter merged into the base weights (W + BA rounded to bf16); not comparable to the un-merged headline"}
    if world == 1 and not args.no_torch_cuda:
        # the same-GPU library baseline (cuBLASLt + SDPA through the oracle port), timed in this run on this box; the full
        # 3 + 10 protocol is `--impl torch_cuda`
        del model, pipe
        torch.cuda.empty_cache()
        try:
            fn, info = torch_cuda_denoise_fn(args, dev, rank)
            tms = _time_cuda(fn, 3, 2, 1, dev)
            tv = args.batch * 3 / (tms / 1000.0)
            line["torch_cuda"] = {"value": tv, "unit": UNIT, "ms_per_step": tms / 3, "steps": 3, "warmup": 2, **info}
            line["vs_torch_cuda"] = value / tv
            del fn
        except Exception as ex:
            line["torch_cuda"] = {"unavailable": f"{type(ex).__name__}: {ex}"}
            line["vs_torch_cuda"] = None
        torch.cuda.empty_cache()
    if world == 1 and not args.no_cpu_baseline and not qwen:
        threads = os.cpu_count() or 1
        cb = cpu_sample_images_per_sec(args.px, args.nfe, threads)
        line["cpu_baseline"] = {"value": cb["value"], "unit": UNIT, "cores": threads, "kind": "port", "extrapolated": True,
                                "cpu_model": _cpu_model_name(), "sample": cb["sample"]}
    _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


# =================================================================================================
# train mode: BASELINE.json configs[3] — one FLUX trajectory-distillation iteration per step
# =================================================================================================
TRAIN_METRIC = "train_samples_per_sec_flux_distill"
TRAIN_UNIT = "samples/s"


def train_flops_per_sample(S_img: int = 4096) -> dict:
    """SURVEY.md §8d, per sample per iteration, no recompute counted: 8 teacher forwards (no LoRA / ArcFlow heads),
    2 student forwards, 2 student backwards (dX chain through the frozen GEMMs + 2.5x attention + 2x LoRA)."""
    f = flux_flops_per_image_nfe(S_img)
    teacher = f["linear"] + f["attn"] + f["embed"]
    student = f["total"]
    bwd = f["linear"] + 2.5 * f["attn"] + 2 * f["lora"] + 2 * f["heads"]
    return dict(teacher_fwd=8 * teacher, student_fwd=2 * student, student_bwd=2 * bwd,
                total=8 * teacher + 2 * student + 2 * bwd)


def run_train(args):
    import torch
    import torch.distributed as dist
    from arcflow_b200 import _lib, build
    from arcflow_b200.config import flux_dev
    from arcflow_b200.model import ArcFluxEngineModel, FluxTeacherEngine
    from arcflow_b200.synthetic import make_flux_inputs, make_flux_state_dict, make_flux_teacher_extras
    from arcflow_b200.train import ArcFlowTrainer, draw_rollout_randoms

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.model != "flux":
        if rank == 0:
            _emit({"mode": "train", "unavailable": "bench.py --mode train times the FLUX configuration (BASELINE.json configs[3])"})
        return 0
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    if rank != 0:
        build.build()
    lib = _lib.load()
    B, px = args.train_batch, args.px
    grid = (px // 16, px // 16)
    cfg = flux_dev()
    sd = make_flux_state_dict(cfg, 1234, dev)
    student = ArcFluxEngineModel(sd, cfg, dev, consume_state_dict=True)
    del sd
    torch.cuda.empty_cache()
    teacher = FluxTeacherEngine(student, make_flux_teacher_extras(cfg, 99, dev))
    student.set_activation_stash("auto")
    trainer = ArcFlowTrainer(student, teacher, state_bits=8)   # configs/flux: AdamW8bit lr 1e-4, betas (.9, .95), clip 50 from iter 100, Karras EMA
    x, txt, pooled = make_flux_inputs(cfg, B, px, px, 512, 42 + rank, dev)     # --diff_seed: per-rank noise / prompts
    g = torch.Generator().manual_seed(rank)
    rands = [draw_rollout_randoms(B, 4, 16, g) for _ in range(2)]
    hx, htxt, hpooled = [t_.cpu().pin_memory() for t_ in (x, txt, pooled)]
    it = [500]       # past the clip / EMA start iterations: the steady-state iteration

    def step_device():
        loss, lv = trainer.train_step(txt, pooled, grid, x, rands, iteration=it[0])
        it[0] += 1
        return loss

    def step_e2e():   # the data-loader hand-off: pinned host batch -> device inside the timed region, loss read back
        loss, lv = trainer.train_step(htxt.to(dev, non_blocking=True), hpooled.to(dev, non_blocking=True), grid,
                                      hx.to(dev, non_blocking=True), rands, iteration=it[0])
        it[0] += 1
        return loss

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        n0 = lib.afb_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for _ in range(steps):
            last = fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), lib.afb_launch_count() - n0, last

    clocks = ClockSampler(local)
    warm = max(args.warmup, 3)
    timed(step_device, 0, warm)
    if rank == 0:
        clocks.start()
    ms_total, launches, loss = timed(step_device, args.steps, 0)
    ms_e2e, _, _ = timed(step_e2e, args.steps, 1)
    clock_info = clocks.stop() if rank == 0 else {}
    identical = None
    if world > 1:   # DDP invariant: every rank holds the same parameters after the all-reduced update
        chk = trainer.opt.params.double().sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        identical = bool((lo == hi).item())
    if rank == 0:
        peaks = _peaks()
        fl = train_flops_per_sample(grid[0] * grid[1])
        step_ms = ms_total / args.steps
        samples = world * B * args.steps
        tf = fl["total"] * B / (step_ms * 1e9)
        _emit({"metric": TRAIN_METRIC, "mode": "train", "value": samples / (ms_total / 1000.0), "unit": TRAIN_UNIT,
               "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": step_ms, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
               "config": {"workload": f"ArcFlow-FLUX trajectory-distillation iteration, batch {B}/GPU, latent 16x{px // 8}x{px // 8} "
                                      f"(S = {512 + grid[0] * grid[1]}): 2 student + 8 teacher forwards, data-free roll-out, adapter-only "
                                      f"backward, flat-arena all-reduce, clip + AdamW8bit (block-wise 8-bit moments) + Karras EMA (BASELINE.json configs[3])",
                          "global_batch": B * world, "parallelism": f"ddp{world}: per-rank noise/prompts, one fp32 all-reduce of "
                                                                    f"the {trainer.opt.n * 4 / 1e9:.2f} GB gradient arena per iteration",
                          "activation_stash": bool(student.activation_stash), "trainable_params": int(trainer.opt.n),
                          "l2": "working set (25 GB weights + activations) >> 126 MB L2; no explicit flush"},
               "e2e": {"value": samples / (ms_e2e / 1000.0), "unit": TRAIN_UNIT,
                       "h2d_bytes_per_step": world * sum(t_.numel() * t_.element_size() for t_ in (hx, htxt, hpooled)),
                       "d2h_bytes_per_step": world * 12},
               "gpu_launches": int(launches), "clocks": clock_info, "loss": loss,
               "roofline": {"kernel": "whole iteration (tcgen05 GEMM + attention forward/backward dominate)", "bound": "tensor",
                            "achieved": tf, "peak": peaks["bf16"], "unit": "TFLOP/s", "frac": tf / peaks["bf16"], "traffic": None,
                            "algorithmic_flops_per_step": fl["total"] * B,
                            "tensor_floor_ms": fl["total"] * B / (peaks["bf16"] * 1e9),
                            "peak_source": f"{peaks['source']} bf16_tflops_sustained"},
               "params_identical_across_ranks": identical,
               "device_mem_used_gb": (torch.cuda.mem_get_info(dev)[1] - torch.cuda.mem_get_info(dev)[0]) / 2 ** 30})
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch_cuda"])
    ap.add_argument("--mode", default="infer", choices=["infer", "train"],
                    help="infer = the headline (2-NFE images/s); train = BASELINE.json configs[3], one distillation iteration")
    ap.add_argument("--train-batch", type=int, default=4, help="--mode train: samples per GPU (the reference's 4)")
    ap.add_argument("--no-torch-cuda", action="store_true", help="skip the same-GPU torch baseline inside our N = 1 run")
    ap.add_argument("--no-cpu-full", action="store_true",
                    help="--impl reference: skip the measured full-depth 256 px loop (BASELINE.json configs[0])")
    ap.add_argument("--model", default="flux", choices=["flux", "qwen"],
                    help="flux = the headline (BASELINE.json configs[1]); qwen = configs[2] (ArcFlow-Qwen-Image, batch-sharded)")
    ap.add_argument("--px", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=8, help="images per GPU per step")
    ap.add_argument("--nfe", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", dest="variants", action="store_false",
                    help="skip the extra (non-headline) fused-adapter measurement at N = 1")
    args = ap.parse_args()
    _claim_stdout()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "torch_cuda":
        return run_torch_cuda(args)
    return run_train(args) if args.mode == "train" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
