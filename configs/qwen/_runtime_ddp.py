# Optimisation / runtime settings shared by the Qwen-Image distillation configs.
# Same keys and values as the reference's configs/qwen/_ddp_train.py (train_cfg clip :14-17, optimizer :18-26,
# lr_config :27-31, runner :32-38); only the keys this build consumes are listed.
train_cfg = dict(diffusion_grad_clip=50.0, diffusion_grad_clip_begin_iter=100)
optimizer = dict(
    diffusion=dict(
        type='AdamW8bit',   # moments are kept in fp32 here: same update rule, no 8-bit state
        lr=1e-4, betas=(0.9, 0.95), weight_decay=0.0,
        paramwise_cfg=dict(custom_keys=dict(proj_out_loggamma=dict(lr_mult=0.1)))))
lr_config = dict(policy='fixed', warmup='linear', warmup_iters=100, warmup_ratio=0.001)
runner = dict(type='DynamicIterBasedRunnerMod', pass_training_status=True, ckpt_trainable_only=True, ckpt_fp16=True,
              ckpt_fp16_ema=True, gc_interval=20)
dist_params = dict(backend='nccl')
module_wrapper = 'ddp'
log_level = 'INFO'
