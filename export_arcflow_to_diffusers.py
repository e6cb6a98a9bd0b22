"""Training checkpoint -> adapter folder that `pipe.load_arcflow_adapter` reads (the reference's
export_arcflow_to_diffusers.py:43-127: `config.json` with `_class_name` + constructor args, and
`diffusion_pytorch_model.safetensors` with the lora_A / lora_B pairs, the three heads and norm_out.linear; EMA weights).

    python export_arcflow_to_diffusers.py configs/flux/arcflux_2nfe_k16.py checkpoints/arcflux_k16_2nfe/latest.pth out_dir
"""
import argparse

from lakonlab.models.builder import student_config
from lakonlab.pipelines.arcflow_loader import write_adapter_folder
from lakonlab.runner import adapter_from_checkpoint, load_checkpoint
from lakonlab.utils import Config


def main(argv=None):
    ap = argparse.ArgumentParser(description='Export an ArcFlow training checkpoint to the diffusers-style adapter folder')
    ap.add_argument('config')
    ap.add_argument('checkpoint')
    ap.add_argument('out_dir')
    ap.add_argument('--no-ema', action='store_true', help='export the raw weights instead of the EMA weights')
    args = ap.parse_args(argv)
    cfg = Config.fromfile(args.config)
    _, model_cfg = student_config(dict(cfg.model.diffusion.denoising))
    adapter = adapter_from_checkpoint(load_checkpoint(args.checkpoint), use_ema=not args.no_ema)
    write_adapter_folder(args.out_dir, model_cfg, adapter)
    print(f'wrote {len(adapter)} tensors to {args.out_dir}')


if __name__ == '__main__':
    main()
