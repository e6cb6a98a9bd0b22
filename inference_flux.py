"""ArcFlow-FLUX few-step inference — the reference's inference_flux.py flow (from_pretrained -> load_arcflow_adapter ->
fixed-shift schedule -> pipe(...)), made runnable offline:

    python inference_flux.py --base /path/to/FLUX.1-dev --adapter /path/to/ArcFlow --subfolder arcflow-flux-2steps \
        --prompt-embeds embeds.pt --nfe 2 --out arcflux_2nfe.pt
    python inference_flux.py --base synthetic://1234 --synthetic-adapter --nfe 2          # no files needed

Text encoders and the VAE are outside this build: prompts enter as cached T5 / CLIP embeddings
(`prompt_embeds` [B, 512, 4096], `pooled_prompt_embeds` [B, 768] in a .pt / .safetensors file), and the result is saved as
latents unless a decoder hook is supplied by the caller.
"""
import argparse

import torch

from lakonlab.pipelines.arcflux_pipeline import ArcFluxPipeline


def _synthetic_adapter_folder(pipe, path, seed=0):
    """Writes a freshly initialised adapter for the loaded base (for dry runs without a trained adapter)."""
    from arcflow_b200.adapter_init import flux_lora_target_paths, init_arcflow_adapter
    from arcflow_b200.config import FLUX_LORA_TARGETS, flux_dev
    from lakonlab.pipelines.arcflow_loader import write_adapter_folder
    cfg = flux_dev()
    base = pipe.transformer.state_dict()
    adapter = init_arcflow_adapter(base, cfg, flux_lora_target_paths(cfg, FLUX_LORA_TARGETS),
                                   generator=torch.Generator().manual_seed(seed))
    write_adapter_folder(path, cfg, adapter)
    return path


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--base', default='synthetic://1234', help="FLUX.1-dev folder (diffusers layout) or synthetic://<seed>")
    ap.add_argument('--adapter', default=None, help='ArcFlow adapter folder (e.g. a local copy of ymyy307/ArcFlow)')
    ap.add_argument('--subfolder', default=None, help="e.g. 'arcflow-flux-2steps'")
    ap.add_argument('--synthetic-adapter', action='store_true', help='initialise an untrained adapter for a dry run')
    ap.add_argument('--prompt-embeds', default=None, help='.pt/.safetensors with prompt_embeds + pooled_prompt_embeds')
    ap.add_argument('--nfe', type=int, default=4)
    ap.add_argument('--width', type=int, default=1024)
    ap.add_argument('--height', type=int, default=1024)
    ap.add_argument('--num-images-per-prompt', type=int, default=1)
    ap.add_argument('--seed', type=int, default=42)
    ap.add_argument('--out', default=None)
    args = ap.parse_args(argv)

    # under `torchrun --nproc-per-node N` the batch (prompts x num_images_per_prompt) is split over the N GPUs inside
    # pipe(...) and the final latents are all-gathered (lakonlab/parallel/batch_parallel.py); rank 0 writes the result
    from lakonlab.parallel import init_from_env
    rank, world, dev = init_from_env()

    pipe = ArcFluxPipeline.from_pretrained(args.base, torch_dtype=torch.bfloat16, device=dev)
    adapter = args.adapter
    if adapter is None:
        if not args.synthetic_adapter:
            raise SystemExit('pass --adapter <folder> (or --synthetic-adapter for a dry run)')
        import tempfile
        adapter = _synthetic_adapter_folder(pipe, tempfile.mkdtemp(prefix='arcflow_adapter_'), args.seed)
    adapter_name = pipe.load_arcflow_adapter(adapter, subfolder=args.subfolder, target_module_name='transformer')
    pipe.scheduler_shift = 3.2   # FlowMatchEulerDiscreteScheduler(shift=3.2, use_dynamic_shifting=False), inference_flux.py:14
    pipe = pipe.to(dev)

    if args.prompt_embeds:
        from lakonlab.datasets import _load_any
        e = _load_any(args.prompt_embeds)
        e = e.get('prompt_embed_kwargs', e)
        prompt_embeds = e.get('prompt_embeds', e.get('encoder_hidden_states'))
        pooled = e.get('pooled_prompt_embeds', e.get('pooled_projections'))
        if prompt_embeds.dim() == 2:
            prompt_embeds, pooled = prompt_embeds[None], pooled[None]
    else:
        g = torch.Generator().manual_seed(args.seed)
        prompt_embeds, pooled = torch.randn(1, 512, 4096, generator=g) * 0.1, torch.randn(1, 768, generator=g)
    out = pipe(prompt_embeds=prompt_embeds.to(dev, torch.bfloat16), pooled_prompt_embeds=pooled.to(dev, torch.bfloat16),
               num_images_per_prompt=args.num_images_per_prompt, width=args.width, height=args.height,
               num_inference_steps=args.nfe, generator=torch.Generator(device=dev).manual_seed(args.seed),
               timestep_ratio=1.0, output_type='latent').images
    path = args.out or f'arcflux_{args.nfe}nfe.pt'
    if rank == 0:
        torch.save(dict(latents=out.cpu(), adapter=adapter_name, nfe=args.nfe, height=args.height, width=args.width), path)
    if rank == 0:
        print(f'{adapter_name}: {tuple(out.shape)} latents -> {path}' + (f' ({world} ranks)' if world > 1 else ''))
    return out


if __name__ == '__main__':
    main()
