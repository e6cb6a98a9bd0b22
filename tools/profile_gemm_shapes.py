"""The named GEMM launch types of one ArcFlow-FLUX forward at the headline size (batch 8, 1024 x 1024: 32768 image rows,
4096 text rows, 36864 joint rows), each launched ALONE with its real operands / epilogue, in a fixed, printed order — so an
`ncu -k regex:gemm_bf16` capture of this script names its launches by position (launch i of the capture = entry i of the
JSON this prints), instead of guessing which launch of a whole step `-s N` landed on.

    python tools/profile_gemm_shapes.py [--reps 5] [--json out.json]          CUDA-event timing (L2 flushed between launches)
    ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -c 8 -o gpurun_out/r02_gemm_shapes \
        python tools/profile_gemm_shapes.py --reps 1 --no-warmup --once

Algorithmic bytes per launch = A (all K-segments) + W + bias + output (+ residual read for the gated epilogues); FLOPs =
2 M N K with the LoRA K-extension counted.
"""
import argparse
import json
import sys

import torch

sys.path.insert(0, ".")
from arcflow_b200 import _lib, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--no-warmup", action="store_true")
ap.add_argument("--once", action="store_true", help="one launch per type, in order (the ncu capture)")
ap.add_argument("--json", default=None)
a = ap.parse_args()

dev = torch.device("cuda", 0)
B, St, Si, D, M, r = 8, 512, 4096, 3072, 12288, 256
S = St + Si
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s, std=1.0: (torch.randn(*s, device=dev, generator=g) * std).bfloat16()

y = rnd(B, S, D)                 # modulated activations (A operand)
attn = rnd(B, S, D)
mlp = rnd(B, S, M, std=0.3)
lt = rnd(B, S, r, std=0.1)
h = rnd(B, S, D)
gate = rnd(B, D, std=0.3)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

img = lambda t: t[:, St:]        # image rows of the joint buffer (strided view, like the engine's)

TYPES = [
    # name, A segments, N, epilogue, uses residual/gate, description
    ("img_qkv", [img(y)], 3 * D, _lib.AFB_EPI_BIAS, False, "double block, image stream QKV: M 32768 x N 9216 x K 3072"),
    ("img_attn_out", [img(attn)], D, _lib.AFB_EPI_BIAS_GATE_RES, True, "double block, image out-projection + gate*y+res: M 32768 x N 3072 x K 3072"),
    ("img_mlp_up", [img(y), img(lt)], M, _lib.AFB_EPI_BIAS_GELU, False, "double block, image MLP-up + LoRA K-ext + GELU: M 32768 x N 12288 x K 3328"),
    ("img_mlp_down", [img(mlp), img(lt)], D, _lib.AFB_EPI_BIAS_GATE_RES, True, "double block, image MLP-down + LoRA K-ext + gate*y+res: M 32768 x N 3072 x K 12544"),
    ("single_qkv", [y], 3 * D, _lib.AFB_EPI_BIAS, False, "single block QKV: M 36864 x N 9216 x K 3072"),
    ("single_mlp_up", [y, lt], M, _lib.AFB_EPI_BIAS_GELU, False, "single block proj_mlp + LoRA + GELU: M 36864 x N 12288 x K 3328"),
    ("single_proj_out", [attn, mlp, lt], D, _lib.AFB_EPI_BIAS_GATE_RES, True, "single block proj_out on [attn | mlp | lora] + gate*y+res: M 36864 x N 3072 x K 15616"),
    ("lora_a_up", [y], r, _lib.AFB_EPI_BIAS, False, "LoRA A-projection of an MLP-up: M 36864 x N 256 x K 3072"),
]

nq = (1 + 0.05 * torch.randn(128, device=dev, generator=g)).bfloat16()
nk = (1 + 0.05 * torch.randn(128, device=dev, generator=g)).bfloat16()
ang = torch.rand(S, 64, device=dev, generator=g) * 50.0
rope = ops.rope_pack(torch.cos(ang).repeat_interleave(2, 1).contiguous(), torch.sin(ang).repeat_interleave(2, 1).contiguous())
TYPES.append(("img_qkv_fused_norm_rope", [img(y)], 3 * D, _lib.AFB_EPI_BIAS, False,
              "image QKV with the fused RMSNorm + RoPE epilogue: M 32768 x N 9216 x K 3072"))

entries = []
for name, segs, N, epi, gated, desc in TYPES:
    K = sum(s.shape[-1] for s in segs)
    rows = segs[0].shape[1]
    w = rnd(N, K, std=0.02)
    bias = rnd(N, std=0.02) if name != "lora_a_up" else None
    if gated:
        out = h[:, St:] if rows == Si else h
        res = out
    else:
        out = torch.empty(B, rows, N, dtype=torch.bfloat16, device=dev)
        res = None
    Mrows = B * rows
    flops = 2.0 * Mrows * N * K
    bytes_alg = 2 * (Mrows * K + N * K + Mrows * N + (Mrows * N if gated else 0) + (N if bias is not None else 0))

    qk = dict(norm_q=nq, norm_k=nk, rope=rope, qk_cols=2 * D, row0=St) if name.endswith("fused_norm_rope") else None

    def launch(segs=segs, w=w, out=out, bias=bias, epi=epi, res=res, qk=qk):
        ops.gemm(segs, w, out, bias=bias, epilogue=epi, gate=gate if res is not None else None, res=res, qk_norm_rope=qk)

    entries.append(dict(name=name, desc=desc, M=Mrows, N=N, K=K, flops=flops, algorithmic_bytes=bytes_alg, launch=launch))

if not a.no_warmup:
    for e in entries:
        e["launch"]()
    torch.cuda.synchronize()

results = []
for e in entries:
    times = []
    for _ in range(1 if a.once else a.reps):
        flush.fill_(1)          # evict the previous launch's output / weights from L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e["launch"]()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = min(times)
    results.append(dict(name=e["name"], desc=e["desc"], M=e["M"], N=e["N"], K=e["K"], ms=ms, ms_all=times,
                        tflops=e["flops"] / (ms * 1e9), algorithmic_bytes=e["algorithmic_bytes"],
                        algorithmic_gbs=e["algorithmic_bytes"] / (ms * 1e6)))
out = dict(what="per-launch-type GEMM timing, isolated, L2 flushed before each launch, CUDA events", batch=B, px=1024,
           order=[e["name"] for e in entries], launches=results)
print(json.dumps(out))
if a.json:
    with open(a.json, "w") as f:
        json.dump(out, f, indent=1)
