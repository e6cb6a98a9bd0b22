#!/usr/bin/env python
"""bench.py — ArcFlow-FLUX 2-NFE 1024x1024 images/sec on N B200s (BASELINE.json metric).

A "step" is one pass of the hot path over one batch: the whole 2-NFE denoise loop (2 transformer
forwards + 2 analytic sampler steps) for 8 images per GPU at 1024x1024 with cached text embeds
(configs[1] of BASELINE.json), synthetic weights of the true FLUX.1-dev + ArcFlow-adapter shapes.

  python bench.py [--gpus N] [--steps K] [--warmup W]           our arm (one rank per GPU under torchrun)
  python bench.py --impl reference [...]                        the reference's CPU torch path (oracle port)

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
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_ncu_traffic.json")
    try:
        with open(path) as f:
            return json.load(f)
    except (OSError, ValueError):
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
    """Times `dbl` double + `sgl` single FLUX blocks (of 19 + 38) of the reference's bf16 torch path for one
    1024^2 image on the host cores and scales by depth and NFE. Embedders/heads/sampler (<0.1 % of the
    FLOPs) are included once, unscaled."""
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
    return dict(value=1.0 / sec_per_image, seconds_sample=dt,
                sample=f"1 image, {dbl} double + {sgl} single of {full[0]}+{full[1]} FLUX blocks at {px}px "
                       f"(S={512 + grid[0] * grid[1]}), bf16 torch CPU, one forward, scaled x{blocks_full / blocks_s:.1f} "
                       f"depth x{nfe} NFE")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if args.model != "flux":
        _emit({"impl": "reference", "unavailable": "the CPU reference arm is implemented for the FLUX headline only"})
        return 0
    import torch
    threads = os.cpu_count() or 1
    for _ in range(args.warmup if args.warmup < 1 else 1):
        cpu_sample_images_per_sec(args.px, args.nfe, threads)
    vals = [cpu_sample_images_per_sec(args.px, args.nfe, threads) for _ in range(max(1, min(args.steps, 3)))]
    best = max(vals, key=lambda d: d["value"])
    line = {
        "impl": "reference", "metric": METRIC, "value": best["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * args.batch / best["value"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": _config(args, 1),
        "cpu_baseline": {"value": best["value"], "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": best["sample"]},
        "e2e": {"value": best["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference deps (diffusers/peft/mmcv) are not installable offline; this is the oracle port of its "
                "torch path (oracle/arcflow_oracle.py) on the host cores",
    }
    _emit(line)
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
        if world > 1:
            dist.all_gather_into_tensor(gathered, r.images)

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
    h2d = sum(t_.numel() * t_.element_size() for t_ in host_in)
    d2h = hout.numel() * hout.element_size()

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
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic", "config": _config(args, world),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "clocks": clock_info,
        "roofline": {
            "kernel": "gemm_bf16_kernel (tcgen05, all Linear/LoRA launches of one step)",
            "bound": "tensor", "achieved": gemm_tf, "peak": peaks["bf16"], "unit": "TFLOP/s",
            "frac": gemm_tf / peaks["bf16"], "traffic": traffic.get("gemm", {}).get("dram_bytes_per_launch"),
            "traffic_launch": traffic.get("gemm", {}).get("launch"), "traffic_source": traffic.get("source"),
            "peak_source": f"{peaks['source']} bf16_tflops_sustained (kernel timed inside a long step)",
            "launches": prof["gemm_launches"], "avg_launch_ms": prof["gemm_ms"] / max(prof["gemm_launches"], 1),
            "algorithmic_flops_per_step": prof["gemm_flops"],
            "share_of_step": prof["gemm_ms"] / step_ms if step_ms else None,
        },
        "roofline_attention": {
            "kernel": "attention_fwd_kernel (tcgen05)", "bound": "tensor", "achieved": attn_tf,
            "peak": peaks["bf16"], "unit": "TFLOP/s", "frac": attn_tf / peaks["bf16"],
            "launches": prof["attn_launches"], "avg_launch_ms": prof["attn_ms"] / max(prof["attn_launches"], 1),
            "algorithmic_bytes_per_launch": 4 * (512 + grid[0] * grid[1]) * 3072 * 2 * args.batch,
            "traffic": traffic.get("attention", {}).get("dram_bytes_per_launch"),
            "share_of_step": prof["attn_ms"] / step_ms if step_ms else None,
        },
        "step_tflops": step_flops / (step_ms * 1e9), "step_frac_of_peak": step_flops / (step_ms * 1e9) / peaks["bf16"],
        "tflop_per_image": fl["total"] * args.nfe / 1e12,
    }
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
        line["variants"] = {"fuse_lora": {"value": args.batch * args.steps / (f0.elapsed_time(f1) / 1000.0), "unit": UNIT,
                                          "note": "adapter merged into the base weights (W + BA rounded to bf16); "
                                                  "not comparable to the un-merged headline"}}
    if world == 1 and not args.no_cpu_baseline and not qwen:
        threads = os.cpu_count() or 1
        cb = cpu_sample_images_per_sec(args.px, args.nfe, threads)
        line["cpu_baseline"] = {"value": cb["value"], "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": cb["sample"]}
    _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
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
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
