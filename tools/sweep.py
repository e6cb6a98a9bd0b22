"""BASELINE.json configs[2] / configs[4]: NFE x resolution sweep (FLUX) and ArcFlow-Qwen at 1024^2 on one GPU.
Prints one JSON line per point: images/s, ms/step, achieved TFLOP/s of algorithmic work, per-kernel-class shares."""
import argparse
import json
import sys

import torch

sys.path.insert(0, ".")
from bench import flux_flops_per_image_nfe, qwen_flops_per_image_nfe, _peaks  # noqa: E402


def run(model_name, px, nfe, batch, txt_len, steps, warmup):
    dev = torch.device("cuda", 0)
    grid = (px // 16, px // 16)
    if model_name == "flux":
        from arcflow_b200.config import flux_dev
        from arcflow_b200.model import ArcFluxEngineModel
        from arcflow_b200.synthetic import make_flux_inputs, make_flux_state_dict
        cfg = flux_dev()
        key = ("flux",)
        if key not in CACHE:
            CACHE.clear()
            sd = make_flux_state_dict(cfg, 1234, dev)
            CACHE[key] = ArcFluxEngineModel(sd, cfg, dev, consume_state_dict=True)
        m = CACHE[key]
        x, txt, pooled = make_flux_inputs(cfg, batch, px, px, txt_len, 42, dev)
        fn = lambda: m.denoise(x, txt, pooled, grid, num_inference_steps=nfe)
        flops = flux_flops_per_image_nfe(grid[0] * grid[1], txt_len)["total"]
    else:
        from arcflow_b200.qwen import ArcQwenEngineModel, make_qwen_inputs, make_qwen_state_dict, qwen_image
        cfg = qwen_image()
        key = ("qwen",)
        if key not in CACHE:
            CACHE.clear()
            torch.cuda.empty_cache()
            sd = make_qwen_state_dict(cfg, 1234, dev)
            CACHE[key] = ArcQwenEngineModel(sd, cfg, dev, consume_state_dict=True)
        m = CACHE[key]
        x, txt = make_qwen_inputs(cfg, batch, px, px, txt_len, 42, dev)
        fn = lambda: m.denoise(x, txt, grid, num_inference_steps=nfe)
        flops = qwen_flops_per_image_nfe(grid[0] * grid[1], txt_len)["total"]
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    m.set_profiling(True); m.read_profile(); fn(); prof = m.read_profile(); m.set_profiling(False)
    tf = flops * nfe * batch / (ms * 1e9)
    return dict(model=model_name, px=px, nfe=nfe, batch=batch, txt_len=txt_len, ms_per_step=ms,
                images_per_s=batch / (ms / 1000), tflops=tf, frac_of_sustained_peak=tf / _peaks()["bf16"],
                gemm_tflops=prof["gemm_flops"] / (prof["gemm_ms"] * 1e9), gemm_share=prof["gemm_ms"] / ms,
                attn_tflops=prof["attn_flops"] / (prof["attn_ms"] * 1e9), attn_share=prof["attn_ms"] / ms,
                attn_gbs_algorithmic=4 * (txt_len + grid[0] * grid[1]) * 3072 * 2 * batch * prof["attn_launches"]
                / (prof["attn_ms"] * 1e6))


CACHE = {}
if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", default="flux:512:2:16,flux:1024:2:8,flux:1024:4:8,flux:1024:8:8,flux:2048:2:2,flux:256:2:8,"
                                        "qwen:1024:2:8")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=2)
    a = ap.parse_args()
    for pt in a.points.split(","):
        name, px, nfe, batch = pt.split(":")
        print(json.dumps(run(name, int(px), int(nfe), int(batch), 512, a.steps, a.warmup)), flush=True)
