# ArcFlow-Qwen-Image (20B), 2 NFE, K = 16, rank-256 adapter: data-free trajectory distillation with a true-CFG teacher.
# Hyper-parameters follow the reference's configs/qwen/arcqwen_2nfe_k16.py (student :24-62, teacher :64-80,
# train_cfg :96-106). `pretrained`: see configs/flux/arcflux_2nfe_k16.py.
_base_ = ['./_data_prompts.py']

name = 'arcqwen_k16_2nfe'
qwen_trunk = dict(in_channels=64, num_layers=60, attention_head_dim=128, num_attention_heads=24, joint_attention_dim=3584,
                  axes_dims_rope=(16, 56, 56), torch_dtype='bfloat16', patch_size=2, freeze=True,
                  pretrained='synthetic://1234')

model = dict(
    type='LatentDiffusionTextImage',
    diffusion=dict(
        type='ArcFlowImitationDataFree',
        policy_type='ArcFlow',
        denoising=dict(
            type='ArcQwenImageTransformer2DModel', num_gaussians=16, logweights_channels=4,
            freeze_exclude=['proj_out_means', 'proj_out_logweights', 'proj_out_loggamma', 'norm_out', 'lora'],
            checkpointing=True, use_lora=True, lora_rank=256, lora_dropout=0.05,   # targets: arcflow_b200.qwen.qwen_lora_targets
            **qwen_trunk),
        flow_loss=dict(type='DiffusionMSELoss', rescale_mode='constant', rescale_cfg=dict(scale=30.0)),
        timestep_sampler=dict(type='ContinuousTimeStepSampler', shift=3.2, logit_normal_enable=False)),
    diffusion_use_ema=True,
    teacher=dict(type='GaussianFlow', denoising=dict(type='QwenImageTransformer2DModel', **qwen_trunk)),
    tie_teacher=True)

train_cfg = dict(num_decay_iters=2000, window_substeps=3, gm_dropout=0.1, num_intermediate_states=4,
                 teacher_guidance_scale=4.0, nfe=2, timestep_ratio=1.0, total_substeps=128)
test_cfg = dict(nfe=2, timestep_ratio=1.0, total_substeps=128)

total_iters = 10000
save_interval = 500
work_dir = f'work_dirs/{name}'
checkpoint_config = dict(interval=save_interval, must_save_interval=1000, by_epoch=False, max_keep_ckpts=1,
                         out_dir='checkpoints/')
log_config = dict(interval=1, hooks=[dict(type='TextLoggerHook')])
custom_hooks = [dict(type='ExponentialMovingAverageHookMod', module_keys=('diffusion_ema',), interp_mode='lerp', interval=1,
                     start_iter=100, momentum_policy='karras', momentum_cfg=dict(gamma=7.0), priority='VERY_HIGH')]
load_from = None
resume_from = f'checkpoints/{name}/latest.pth'
workflow = [('train', save_interval)]

# ---- optimisation / runtime (values of the reference's configs/qwen/_ddp_train.py: clip :14-17, optimizer :18-26,
# lr_config :27-31, runner :32-38; only the keys this build consumes) ----
train_cfg.update(diffusion_grad_clip=50.0, diffusion_grad_clip_begin_iter=100)
optimizer = dict(diffusion=dict(
    type='AdamW8bit',   # block-wise 8-bit moments as in bitsandbytes (arcflow_b200/optim.py); optim_bits=32 keeps fp32 moments
    lr=1e-4, betas=(0.9, 0.95), weight_decay=0.0,
    paramwise_cfg=dict(custom_keys=dict(proj_out_loggamma=dict(lr_mult=0.1)))))
lr_config = dict(policy='fixed', warmup='linear', warmup_iters=100, warmup_ratio=0.001)
runner = dict(type='DynamicIterBasedRunnerMod', pass_training_status=True, ckpt_trainable_only=True, ckpt_fp16=True,
              ckpt_fp16_ema=True, gc_interval=20)
dist_params = dict(backend='nccl')
module_wrapper = 'ddp'
log_level = 'INFO'
