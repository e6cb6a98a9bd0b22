# Depth/width-reduced ArcFlow-FLUX of the true structure on synthetic weights: exercises train.py end to end in seconds
# (tests/test_gpu_entrypoints.py). Not a reference configuration.
_base_ = ['./arcflux_2nfe_k16.py']

name = 'arcflux_tiny_smoke'
flux_trunk = dict(num_layers=2, num_single_layers=2, num_attention_heads=2, joint_attention_dim=256,
                  pooled_projection_dim=256)
model = dict(
    diffusion=dict(denoising=dict(lora_rank=256, **flux_trunk)),
    teacher=dict(denoising=dict(**flux_trunk)))
train_cfg = dict(num_decay_iters=4, diffusion_grad_clip_begin_iter=1)
lr_config = dict(warmup_iters=2)
data = dict(
    train_dataloader=dict(samples_per_gpu=2),
    train=dict(type='SyntheticPrompts', joint_attention_dim=256, pooled_projection_dim=256, seq_len=64,
               latent_size=(16, 16, 16), length=64))
total_iters = 4
save_interval = 2
work_dir = f'work_dirs/{name}'
checkpoint_config = dict(interval=save_interval, max_keep_ckpts=2, out_dir='work_dirs/checkpoints/')
custom_hooks = [dict(type='ExponentialMovingAverageHookMod', module_keys=('diffusion_ema',), interp_mode='lerp', interval=1,
                     start_iter=1, momentum_policy='karras', momentum_cfg=dict(gamma=7.0), priority='VERY_HIGH')]
resume_from = f'work_dirs/checkpoints/{name}/latest.pth'
workflow = [('train', save_interval)]
