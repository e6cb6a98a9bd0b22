# Prompt-embedding source for data-free training (the role of the reference's configs/flux/_data_trainval.py).
# Point `cache_dir` at a folder of cached T5/CLIP embeddings (*.pt / *.safetensors holding prompt_embed_kwargs), or keep
# the synthetic source for benchmarking.
data = dict(
    workers_per_gpu=0,
    train_dataloader=dict(samples_per_gpu=4),
    train=dict(type='SyntheticPrompts', joint_attention_dim=4096, pooled_projection_dim=768, seq_len=512,
               latent_size=(16, 128, 128)),
    # train=dict(type='ImagePrompts', cache_dir='data/prompt_cache/flux', pad_seq_len=512, latent_size=(16, 128, 128)),
)
