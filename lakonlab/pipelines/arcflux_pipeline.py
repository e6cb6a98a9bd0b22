"""ArcFluxPipeline — the reference's text-to-image pipeline surface for the 2-NFE ArcFlow sampler,
re-built on the native engine.

Keeps the call signature and semantics of `ArcFluxPipeline.__call__`
(lakonlab/pipelines/arcflux_pipeline.py:252-276, loop :453-524): fp32 packed latents, the timestep schedule of
`retrieve_raw_timesteps` + fixed shift, one transformer call and one analytic momentum-integration step per
NFE, `FluxPipelineOutput(images=...)`. Out of scope here (metric uses cached embeds and latent
output): the CLIP/T5 text encoders and the VAE — `prompt=` needs a `text_encoder_fn`, `output_type='pil'`
needs a `vae_decode_fn`, both optional hooks.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional, Union

import torch

from arcflow_b200 import ops
from arcflow_b200.schedule import denoise_sigmas, retrieve_raw_timesteps  # noqa: F401  (same public name)
from lakonlab.parallel.batch_parallel import gather_latents, parallel_context, shard_bounds
from .arcflow_loader import ArcFlowLoaderMixin


@dataclass
class FluxPipelineOutput:
    images: Any


class FluxBaseTransformer:
    """Holder of a stock FLUX transformer state dict until `load_arcflow_adapter` swaps it."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda"):
        self._sd = state_dict
        self.device = torch.device(device)
        self.dtype = torch.bfloat16

    def state_dict(self):
        return self._sd


def _load_base_transformer(path: str, arch: str, device):
    import os
    if path.startswith("synthetic://"):
        from lakonlab.models.builder import synthetic_base_state_dict
        if arch == "flux":
            from arcflow_b200.config import flux_dev as full_cfg
        else:
            from arcflow_b200.qwen import qwen_image as full_cfg
        return synthetic_base_state_dict(arch, full_cfg(), int(path[len("synthetic://"):] or 1234), device)
    from lakonlab.models.builder import load_transformer_weights
    sub = os.path.join(path, "transformer")
    return load_transformer_weights(sub if os.path.isdir(sub) else path)


class ArcFluxPipeline(ArcFlowLoaderMixin):
    vae_scale_factor = 8
    default_sample_size = 128

    def __init__(self, transformer=None, scheduler_shift: float = 3.2, text_encoder_fn: Optional[Callable] = None,
                 vae_decode_fn: Optional[Callable] = None, policy_type: str = "ArcFlow", vae=None):
        if policy_type != "ArcFlow":
            raise ValueError(f"Invalid policy: {policy_type}. Supported policies are ['ArcFlow'].")
        self.transformer = transformer
        self.scheduler_shift = scheduler_shift
        self.text_encoder_fn = text_encoder_fn
        self.vae_decode_fn = vae_decode_fn
        self.vae = vae      # arcflow_b200.vae.FluxVAEDecoder (native decode, arcflux_pipeline.py:531-534 of the reference)
        self._num_timesteps = 0
        self._interrupt = False
        self.use_cuda_graph = False
        self.batch_parallel = True

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, torch_dtype=torch.bfloat16, device="cuda", **kwargs):
        """`ArcFluxPipeline.from_pretrained('black-forest-labs/FLUX.1-dev', torch_dtype=bf16)` (inference_flux.py:5-7),
        offline: a local diffusers-layout folder (its `transformer/` sub-folder, or the folder itself, holds the
        `.safetensors` shards) or `synthetic://<seed>` for seeded weights of the FLUX.1-dev shape. Text encoders / VAE
        are attached by the caller as `text_encoder_fn` / `vae_decode_fn` hooks."""
        if torch_dtype not in (None, torch.bfloat16):
            raise ValueError("this build computes in bf16 only")
        return cls(transformer=FluxBaseTransformer(_load_base_transformer(pretrained_model_name_or_path, "flux", device),
                                                   device=device), **kwargs)

    def enable_cuda_graph(self, on: bool = True):
        """Extension (not in the reference): replay the denoising loop as ONE captured CUDA graph per (shape, schedule)
        instead of ~1300 kernel launches — for small batches / resolutions where the loop is launch-bound."""
        self.use_cuda_graph = bool(on)
        return self

    def enable_batch_parallel(self, on: bool = True):
        """Extension (the reference's README.md:39 To-Do): when `torch.distributed` is initialised, one `pipe(...)` call
        made with the SAME arguments on every rank splits the batch over the ranks (weights replicated, nothing exchanged
        inside the loop) and all-gathers the final packed latents, so every rank returns the full batch. On by default;
        off = every rank denoises whatever it was given."""
        self.batch_parallel = bool(on)
        return self

    @property
    def num_timesteps(self):
        return self._num_timesteps

    def fuse_lora(self, **_unused):
        """diffusers' `pipe.fuse_lora()`: merge the ArcFlow adapter's low-rank branches into the base weights of
        `pipe.transformer` (ArcFluxEngineModel.fuse_lora) — same images up to bf16 rounding of the merged weights, ~5 % less
        work per step. Inference only, one-way."""
        self.transformer.fuse_lora()
        return self

    @property
    def interrupt(self):
        return self._interrupt

    def to(self, device=None, *args, **kwargs):
        """The engine packs its weights on the device given at load time; moving an engine-backed transformer is not
        supported (there is no CPU path), so anything but that device is an error rather than a silent no-op."""
        tr = self.transformer
        if device is not None and tr is not None and hasattr(tr, "device") and not isinstance(device, torch.dtype):
            want = torch.device(device)
            have = torch.device(tr.device)
            if want.type != have.type or (want.index is not None and have.index is not None and want.index != have.index):
                raise RuntimeError(f"pipe.to({device!r}): the transformer lives on {have}; load the pipeline with device=")
        return self

    # -- layout helpers kept under the reference's names (arcflux_pipeline.py:163-193) ---------------
    @staticmethod
    def _pack_latents(latents, batch_size, num_channels_latents, height, width, patch_size=1, target_patch_size=2):
        s = target_patch_size // patch_size
        x = latents.view(batch_size, num_channels_latents * patch_size * patch_size, height // target_patch_size, s,
                         width // target_patch_size, s)
        return x.permute(0, 2, 4, 1, 3, 5).reshape(batch_size, (height // target_patch_size) * (width // target_patch_size),
                                                   num_channels_latents * target_patch_size * target_patch_size)

    @staticmethod
    def _unpack_latents(latents, height, width, vae_scale_factor, patch_size=2, target_patch_size=1):
        b, _, ch = latents.shape
        s = patch_size // target_patch_size
        h, w = int(height) // (vae_scale_factor * patch_size), int(width) // (vae_scale_factor * patch_size)
        x = latents.view(b, h, w, ch // (s * s), s, s).permute(0, 3, 1, 4, 2, 5)
        return x.reshape(b, ch // (s * s), h * s, w * s)

    @staticmethod
    def postprocess_image(image: torch.Tensor, output_type: str = "pil"):
        """diffusers VaeImageProcessor.postprocess: denormalise to [0, 1]; 'pt' -> fp32 [B, 3, H, W]; 'np' -> [B, H, W, 3];
        'pil' -> list of PIL images."""
        image = (image / 2 + 0.5).clamp(0, 1)
        if output_type == "pt":
            return image
        arr = image.permute(0, 2, 3, 1).float().cpu().numpy()
        if output_type == "np":
            return arr
        if output_type != "pil":
            raise ValueError(f"unsupported output_type '{output_type}' (latent, pt, np, pil)")
        from PIL import Image
        return [Image.fromarray((a * 255).round().astype("uint8")) for a in arr]

    def prepare_latents(self, batch_size, num_channels_latents, height, width, dtype, device, generator, latents=None):
        h, w = 2 * (int(height) // (self.vae_scale_factor * 2)), 2 * (int(width) // (self.vae_scale_factor * 2))
        if latents is not None:
            return latents.to(device=device, dtype=dtype, non_blocking=True)
        shape = (batch_size, num_channels_latents, h, w)
        if isinstance(generator, list):
            noise = torch.cat([torch.randn((1, *shape[1:]), generator=g, device=g.device, dtype=dtype).to(device)
                               for g in generator], 0)
        else:
            gdev = generator.device if generator is not None else device
            noise = torch.randn(shape, generator=generator, device=gdev, dtype=dtype).to(device)
        return self._pack_latents(noise, batch_size, num_channels_latents, h, w)

    @torch.inference_mode()
    def __call__(self, prompt: Union[str, List[str]] = None, prompt_2=None, height: Optional[int] = None,
                 width: Optional[int] = None, num_inference_steps: int = 4, total_substeps: int = 128,
                 timestep_ratio: float = 0.5, temperature: Union[float, str] = "auto", guidance_scale: float = 3.5,
                 num_images_per_prompt: Optional[int] = 1, generator=None, latents: Optional[torch.Tensor] = None,
                 prompt_embeds: Optional[torch.Tensor] = None, pooled_prompt_embeds: Optional[torch.Tensor] = None,
                 output_type: Optional[str] = "pil", return_dict: bool = True,
                 joint_attention_kwargs: Optional[Dict[str, Any]] = None,
                 callback_on_step_end: Optional[Callable[[Any, int, Any, Dict], Dict]] = None,
                 callback_on_step_end_tensor_inputs: List[str] = ["latents"], max_sequence_length: int = 512,
                 ip_adapter_image=None, ip_adapter_image_embeds=None):
        if ip_adapter_image is not None or ip_adapter_image_embeds is not None:
            # accepted for signature compatibility (arcflux_pipeline.py:268-269); the reference forwards them to the stock
            # FluxPipeline image-projection path, which is outside the ArcFlow hot path
            raise NotImplementedError("IP-Adapter inputs are out of scope of this build (SURVEY.md §8)")
        tr = self.transformer
        if tr is None or not hasattr(tr, "denoise"):
            raise RuntimeError("pipe.transformer is not an ArcFlow module — call pipe.load_arcflow_adapter(...) first")
        height = height or self.default_sample_size * self.vae_scale_factor
        width = width or self.default_sample_size * self.vae_scale_factor
        if height % 16 or width % 16:
            raise ValueError(f"`height` and `width` have to be divisible by 16 but are {height} and {width}.")
        # diffusers' `joint_attention_kwargs={"scale": s}` = runtime LoRA scale (scale_lora_layers, arcflux.py:150-156)
        tr.set_lora_scale(float((joint_attention_kwargs or {}).get("scale", 1.0)))
        device = tr.device
        if prompt_embeds is None:
            if prompt is None:
                raise ValueError("Provide either `prompt` or `prompt_embeds`.")
            if self.text_encoder_fn is None:
                raise NotImplementedError("text encoders are out of scope of this build: pass cached `prompt_embeds` "
                                          "and `pooled_prompt_embeds`, or construct the pipeline with text_encoder_fn")
            prompt_embeds, pooled_prompt_embeds = self.text_encoder_fn(prompt, max_sequence_length)
        if pooled_prompt_embeds is None:
            raise ValueError("If `prompt_embeds` are provided, `pooled_prompt_embeds` also have to be passed.")
        if num_images_per_prompt and num_images_per_prompt > 1:
            prompt_embeds = prompt_embeds.repeat_interleave(num_images_per_prompt, 0)
            pooled_prompt_embeds = pooled_prompt_embeds.repeat_interleave(num_images_per_prompt, 0)
        batch = prompt_embeds.shape[0]
        # batch-parallel call: this rank keeps images [lo, hi) — sliced on the HOST side of the copy, and
        # noise is drawn for the whole batch from the caller's generator, so the result does not depend on the rank count
        rank, world = parallel_context(self.batch_parallel)
        lo, hi = shard_bounds(batch, rank, world)
        num_channels_latents = tr.cfg.in_channels // 4
        if latents is not None:
            latents = latents[lo:hi]
        elif world > 1:
            if isinstance(generator, list):
                generator = generator[lo:hi]
                latents = self.prepare_latents(hi - lo, num_channels_latents, height, width, torch.float32, device, generator)
            else:
                gdev = generator.device if generator is not None else torch.device("cpu")
                latents = self.prepare_latents(batch, num_channels_latents, height, width, torch.float32, gdev, generator)[lo:hi]
        prompt_embeds = prompt_embeds[lo:hi].to(device, non_blocking=True)
        pooled_prompt_embeds = pooled_prompt_embeds[lo:hi].to(device, non_blocking=True)
        latents = self.prepare_latents(hi - lo, num_channels_latents, height, width, torch.float32, device, generator, latents)
        grid = (height // 16, width // 16)
        _, _, total = retrieve_raw_timesteps(num_inference_steps, total_substeps, timestep_ratio)
        self._num_timesteps = total

        if hi == lo:
            pass    # more ranks than images: this rank only takes part in the gather
        elif callback_on_step_end is None:
            # whole loop in one C-ABI call (transformer + sampler per NFE, no host round trips)
            latents = tr.denoise(latents, prompt_embeds, pooled_prompt_embeds, grid, num_inference_steps=num_inference_steps,
                                 total_substeps=total_substeps, timestep_ratio=timestep_ratio, shift=self.scheduler_shift,
                                 guidance_scale=guidance_scale, cuda_graph=self.use_cuda_graph)
        else:
            sig = denoise_sigmas(num_inference_steps, total_substeps, timestep_ratio, self.scheduler_shift)
            for i in range(num_inference_steps):
                if self.interrupt:
                    continue
                head = tr.forward_heads(latents, prompt_embeds, pooled_prompt_embeds, sig[i], guidance_scale, grid)
                latents = ops.sampler_step(head.reshape(-1, head.shape[-1]), latents, sig[i], sig[i], sig[i + 1],
                                           num_gaussians=tr.num_gaussians)
                t_src = torch.tensor(sig[i] * 1000.0, device=device)
                tensors = dict(latents=latents, prompt_embeds=prompt_embeds, pooled_prompt_embeds=pooled_prompt_embeds)
                cb = callback_on_step_end(self, i, t_src, {k: tensors[k] for k in callback_on_step_end_tensor_inputs})
                latents = cb.pop("latents", latents)
                prompt_embeds = cb.pop("prompt_embeds", prompt_embeds)

        if world > 1:   # the sampler-boundary exchange: one all-gather of the final packed latents (1.05 MB / image)
            latents = gather_latents(latents.contiguous(), batch)
        if output_type == "latent":
            image = latents
        elif self.vae is not None:
            # latents / scaling_factor + shift_factor -> vae.decode -> image_processor.postprocess (reference :531-534);
            # the affine map runs inside the decoder's first kernel
            image = self.postprocess_image(self.vae.decode(self._unpack_latents(latents, height, width, self.vae_scale_factor)),
                                           output_type)
        else:
            if self.vae_decode_fn is None:
                raise NotImplementedError("no VAE attached: use output_type='latent', or construct the pipeline with "
                                          "vae=arcflow_b200.vae.FluxVAEDecoder(...) or a vae_decode_fn hook")
            image = self.vae_decode_fn(self._unpack_latents(latents, height, width, self.vae_scale_factor), output_type)
        if not return_dict:
            return (image,)
        return FluxPipelineOutput(images=image)
