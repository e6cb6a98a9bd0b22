"""ArcQwenImagePipeline — Qwen-Image variant of the ArcFlow sampler surface on the native engine.

Keeps the call contract of `ArcQwenImagePipeline.__call__` (lakonlab/pipelines/arcqwen_pipeline.py:239-259, loop
:395-463): variable-length text via `prompt_embeds` + `prompt_embeds_mask` (the wrapper trims to the longest
prompt of the batch, arcqwen.py:325-330; the mask is not used inside attention in diffusers 0.35.1), no
guidance embedding, fp32 packed latents. Text encoder and VAE are out of scope (hooks).
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Union

import torch

from arcflow_b200 import ops
from arcflow_b200.schedule import denoise_sigmas, retrieve_raw_timesteps
from lakonlab.parallel.batch_parallel import gather_latents, parallel_context, shard_bounds
from .arcflow_loader import ArcFlowLoaderMixin
from .arcflux_pipeline import ArcFluxPipeline, FluxPipelineOutput


class QwenImagePipelineOutput(FluxPipelineOutput):
    pass


class ArcQwenImagePipeline(ArcFlowLoaderMixin):
    vae_scale_factor = 8
    default_sample_size = 128
    _pack_latents = staticmethod(ArcFluxPipeline._pack_latents)
    _unpack_latents = staticmethod(ArcFluxPipeline._unpack_latents)
    enable_batch_parallel = ArcFluxPipeline.enable_batch_parallel
    to = ArcFluxPipeline.to

    def __init__(self, transformer=None, scheduler_shift: float = 3.2, text_encoder_fn: Optional[Callable] = None,
                 vae_decode_fn: Optional[Callable] = None, policy_type: str = "ArcFlow"):
        if policy_type != "ArcFlow":
            raise ValueError(f"Invalid policy: {policy_type}. Supported policies are ['ArcFlow'].")
        self.transformer = transformer
        self.scheduler_shift = scheduler_shift
        self.text_encoder_fn = text_encoder_fn
        self.vae_decode_fn = vae_decode_fn
        self._num_timesteps = 0
        self._interrupt = False
        self.use_cuda_graph = False
        self.batch_parallel = True

    def enable_cuda_graph(self, on: bool = True):
        """Extension (not in the reference): replay the denoising loop as one captured CUDA graph per (shape, schedule)."""
        self.use_cuda_graph = bool(on)
        return self

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, torch_dtype=torch.bfloat16, device="cuda", **kwargs):
        """`ArcQwenImagePipeline.from_pretrained('Qwen/Qwen-Image', torch_dtype=bf16)` (inference_qwen.py:5-7), offline:
        see ArcFluxPipeline.from_pretrained."""
        from .arcflux_pipeline import FluxBaseTransformer, _load_base_transformer
        if torch_dtype not in (None, torch.bfloat16):
            raise ValueError("this build computes in bf16 only")
        return cls(transformer=FluxBaseTransformer(_load_base_transformer(pretrained_model_name_or_path, "qwen", device),
                                                   device=device), **kwargs)

    def fuse_lora(self, **_unused):
        """diffusers' `pipe.fuse_lora()`: merge the ArcFlow adapter's low-rank branches into the base weights of
        `pipe.transformer` (ArcFluxEngineModel.fuse_lora) — same images up to bf16 rounding of the merged weights, ~5 % less
        work per step. Inference only, one-way."""
        self.transformer.fuse_lora()
        return self

    @property
    def interrupt(self):
        return self._interrupt

    @torch.inference_mode()
    def __call__(self, prompt: Union[str, List[str]] = None, height: Optional[int] = None, width: Optional[int] = None,
                 num_inference_steps: int = 4, total_substeps: int = 128, timestep_ratio: float = 0.5,
                 num_images_per_prompt: int = 1, generator=None, latents: Optional[torch.Tensor] = None,
                 prompt_embeds: Optional[torch.Tensor] = None, prompt_embeds_mask: Optional[torch.Tensor] = None,
                 output_type: Optional[str] = "pil", return_dict: bool = True,
                 attention_kwargs: Optional[Dict[str, Any]] = None,
                 callback_on_step_end: Optional[Callable] = None,
                 callback_on_step_end_tensor_inputs: List[str] = ["latents"], max_sequence_length: int = 512):
        tr = self.transformer
        if tr is None or not hasattr(tr, "denoise"):
            raise RuntimeError("pipe.transformer is not an ArcFlow module — call pipe.load_arcflow_adapter(...) first")
        height = height or self.default_sample_size * self.vae_scale_factor
        width = width or self.default_sample_size * self.vae_scale_factor
        if height % 16 or width % 16:
            raise ValueError(f"`height` and `width` have to be divisible by 16 but are {height} and {width}.")
        self.transformer.set_lora_scale(float((attention_kwargs or {}).get("scale", 1.0)))   # runtime LoRA scale
        device = tr.device
        if prompt_embeds is None:
            if prompt is None:
                raise ValueError("Provide either `prompt` or `prompt_embeds`.")
            if self.text_encoder_fn is None:
                raise NotImplementedError("the text encoder is out of scope of this build: pass cached `prompt_embeds` "
                                          "(+ `prompt_embeds_mask`) or construct the pipeline with text_encoder_fn")
            prompt_embeds, prompt_embeds_mask = self.text_encoder_fn(prompt, max_sequence_length)
        if prompt_embeds_mask is not None:   # trim to the longest prompt (arcqwen.py:325-330)
            max_len = int(prompt_embeds_mask.sum(dim=1).max().item())
            prompt_embeds = prompt_embeds[:, :max_len]
        if num_images_per_prompt > 1:
            prompt_embeds = prompt_embeds.repeat_interleave(num_images_per_prompt, 0)
        batch = prompt_embeds.shape[0]
        # batch-parallel call: this rank keeps images [lo, hi); noise is drawn for the whole batch from
        # the caller's generator so the result does not depend on the rank count
        rank, world = parallel_context(self.batch_parallel)
        lo, hi = shard_bounds(batch, rank, world)
        prompt_embeds = prompt_embeds[lo:hi].to(device, non_blocking=True)
        h, w = 2 * (height // 16), 2 * (width // 16)
        if latents is None:
            gdev = generator.device if generator is not None else (device if world == 1 else torch.device("cpu"))
            noise = torch.randn((batch, 16, h, w), generator=generator, device=gdev, dtype=torch.float32)[lo:hi].to(device)
            latents = self._pack_latents(noise, hi - lo, 16, h, w)
        else:
            latents = latents[lo:hi].to(device=device, dtype=torch.float32, non_blocking=True)
        grid = (height // 16, width // 16)
        self._num_timesteps = retrieve_raw_timesteps(num_inference_steps, total_substeps, timestep_ratio)[2]
        if hi == lo:
            pass    # more ranks than images: this rank only takes part in the gather
        elif callback_on_step_end is None:
            latents = tr.denoise(latents, prompt_embeds, grid, num_inference_steps=num_inference_steps,
                                 total_substeps=total_substeps, timestep_ratio=timestep_ratio, shift=self.scheduler_shift,
                                 cuda_graph=self.use_cuda_graph)
        else:
            sig = denoise_sigmas(num_inference_steps, total_substeps, timestep_ratio, self.scheduler_shift)
            for i in range(num_inference_steps):
                if self.interrupt:
                    continue
                head = tr.forward_heads(latents, prompt_embeds, sig[i], grid)
                latents = ops.sampler_step(head.reshape(-1, head.shape[-1]), latents, sig[i], sig[i], sig[i + 1],
                                           num_gaussians=tr.num_gaussians)
                tensors = dict(latents=latents, prompt_embeds=prompt_embeds)
                cb = callback_on_step_end(self, i, torch.tensor(sig[i] * 1000.0, device=device),
                                          {k: tensors[k] for k in callback_on_step_end_tensor_inputs})
                latents = cb.pop("latents", latents)
                prompt_embeds = cb.pop("prompt_embeds", prompt_embeds)
        if world > 1:   # the sampler-boundary exchange: one all-gather of the final packed latents
            latents = gather_latents(latents.contiguous(), batch)
        if output_type == "latent":
            image = latents
        else:
            if self.vae_decode_fn is None:
                raise NotImplementedError("the VAE is out of scope of this build: use output_type='latent' "
                                          "or construct the pipeline with vae_decode_fn")
            image = self.vae_decode_fn(self._unpack_latents(latents, height, width, self.vae_scale_factor)[:, :, None],
                                       output_type)
        return QwenImagePipelineOutput(images=image) if return_dict else (image,)
