"""ArcFlow-Qwen-Image few-step inference — the reference's inference_qwen.py flow (from_pretrained -> load_arcflow_adapter ->
fixed-shift schedule -> pipe(...)), made runnable offline:

    python inference_qwen.py --base /path/to/Qwen-Image --adapter /path/to/ArcFlow --subfolder arcflow-qwen-2steps \
        --prompt-embeds embeds.pt --nfe 2 --out arcqwen_2nfe.pt
    python inference_qwen.py --base synthetic://1234 --synthetic-adapter --nfe 2          # no files needed

Text encoders and the VAE are outside this build: prompts enter as cached Qwen2.5-VL embeddings
(`prompt_embeds` [B, S_t, 3584] (+ optional `prompt_embeds_mask`) in a .pt / .safetensors file), and the result is saved as
latents unless a decoder hook is supplied by the caller.
"""
import argparse

import torch

from lakonlab.pipelines.arcqwen_pipeline import ArcQwenImagePipeline


def _synthetic_adapter_folder(pipe, path, seed=0):
    """Writes a freshly initialised adapter for the loaded base (for dry runs without a trained adapter)."""
    from arcflow_b200.adapter_init import init_arcflow_adapter
    from arcflow_b200.qwen import qwen_image, qwen_lora_targets
    from lakonlab.pipelines.arcflow_loader import write_adapter_folder
    cfg = qwen_image()
    base = pipe.transformer.state_dict()
    adapter = init_arcflow_adapter(base, cfg,
                                   qwen_lora_targets(cfg), generator=torch.Generator().manual_seed(seed))
    write_adapter_folder(path, cfg, adapter)
    return path


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--base', default='synthetic://1234', help="Qwen-Image folder (diffusers layout) or synthetic://<seed>")
    ap.add_argument('--adapter', default=None, help='ArcFlow adapter folder (e.g. a local copy of ymyy307/ArcFlow)')
    ap.add_argument('--subfolder', default=None, help="e.g. 'arcflow-qwen-2steps'")
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

    pipe = ArcQwenImagePipeline.from_pretrained(args.base, torch_dtype=torch.bfloat16, device=dev)
    adapter = args.adapter
    if adapter is None:
        if not args.synthetic_adapter:
            raise SystemExit('pass --adapter <folder> (or --synthetic-adapter for a dry run)')
        import tempfile
        adapter = _synthetic_adapter_folder(pipe, tempfile.mkdtemp(prefix='arcflow_adapter_'), args.seed)
    adapter_name = pipe.load_arcflow_adapter(adapter, subfolder=args.subfolder, target_module_name='transformer')
    pipe.scheduler_shift = 3.2   # FlowMatchEulerDiscreteScheduler(shift=3.2, use_dynamic_shifting=False), inference_qwen.py:14
    pipe = pipe.to(dev)

    mask = None
    if args.prompt_embeds:
        from lakonlab.datasets import _load_any
        e = _load_any(args.prompt_embeds)
        e = e.get('prompt_embed_kwargs', e)
        prompt_embeds = e.get('prompt_embeds', e.get('encoder_hidden_states'))
        mask = e.get('prompt_embeds_mask', e.get('encoder_hidden_states_mask'))
        if prompt_embeds.dim() == 2:
            prompt_embeds = prompt_embeds[None]
            mask = mask[None] if mask is not None else None
    else:
        g = torch.Generator().manual_seed(args.seed)
        prompt_embeds = torch.randn(1, 512, 3584, generator=g)
    out = pipe(prompt_embeds=prompt_embeds.to(dev, torch.bfloat16),
               prompt_embeds_mask=mask.to(dev) if mask is not None else None,
               num_images_per_prompt=args.num_images_per_prompt, width=args.width, height=args.height,
               num_inference_steps=args.nfe, generator=torch.Generator(device=dev).manual_seed(args.seed),
               timestep_ratio=1.0, output_type='latent').images
    path = args.out or f'arcqwen_{args.nfe}nfe.pt'
    if rank == 0:
        torch.save(dict(latents=out.cpu(), adapter=adapter_name, nfe=args.nfe, height=args.height, width=args.width), path)
    if rank == 0:
        print(f'{adapter_name}: {tuple(out.shape)} latents -> {path}' + (f' ({world} ranks)' if world > 1 else ''))
    return out


if __name__ == '__main__':
    main()
