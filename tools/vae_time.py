"""Times the native FLUX VAE decoder at the headline size (1024 x 1024: 128 x 128 latents) next to the original BFL
autoencoder run by PyTorch's own CUDA libraries (cuDNN convolutions + SDPA) in bf16 on the same GPU, and checks the two
against each other. Usage: python tools/vae_time.py [batch] [latent_h] [latent_w]"""
import json
import sys

import torch

sys.path.insert(0, ".")
from arcflow_b200 import _lib  # noqa: E402
from arcflow_b200.vae import FluxVAEDecoder  # noqa: E402
from oracle import vae_oracle as V  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
h = int(sys.argv[2]) if len(sys.argv) > 2 else 128
w = int(sys.argv[3]) if len(sys.argv) > 3 else h
dev = torch.device("cuda", 0)
sd = V.make_vae_decoder_state_dict(seed=7, device=dev)
z = torch.randn(B, 16, h, w, device=dev, generator=torch.Generator(device=dev).manual_seed(1)) * 0.8
dec = FluxVAEDecoder(sd, device=dev)
lib = _lib.load()


def timed(fn, n=5, warm=2):
    for _ in range(warm):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = lib.afb_launch_count()
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (lib.afb_launch_count() - n0) // n, out


ms, launches, img = timed(lambda: dec.decode(z))
fl = FluxVAEDecoder.flops(h, w) * B
res = dict(what="FLUX VAE decode", batch=B, latent=[h, w], image=[8 * h, 8 * w], ms=ms, images_per_s=B / (ms / 1e3),
           tflops=fl / (ms * 1e9), tflop_per_image=fl / B / 1e12, launches=launches,
           mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30)
try:
    with torch.no_grad():
        tms, _, ref = timed(lambda: V.vae_decode(sd, z, dtype=torch.bfloat16), n=3, warm=1)
    res["torch_cuda_bf16"] = dict(ms=tms, images_per_s=B / (tms / 1e3), tflops=fl / (tms * 1e9),
                                  kernels="torch.nn.functional.conv2d (cuDNN) + group_norm + SDPA, NCHW bf16")
    res["vs_torch_cuda"] = tms / ms
    res["rel_l2_vs_torch_cuda_bf16"] = ((img - ref.float()).norm() / ref.float().norm()).item()
    with torch.no_grad():
        ref32 = V.vae_decode(sd, z[:1], dtype=torch.float32)
    res["rel_l2_vs_fp32_first_image"] = ((img[:1] - ref32).norm() / ref32.norm()).item()
    res["torch_bf16_rel_l2_vs_fp32_first_image"] = ((ref[:1].float() - ref32).norm() / ref32.norm()).item()
except Exception as ex:   # the baseline is optional; the native figure stands
    res["torch_cuda_bf16"] = dict(unavailable=f"{type(ex).__name__}: {ex}")
print(json.dumps(res))
