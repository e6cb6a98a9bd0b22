# Prompt-embedding source for data-free training (the role of the reference's configs/qwen/_data_trainval.py).
# The Qwen teacher uses true classifier-free guidance, so every sample also carries the negative-prompt embedding.
data = dict(
    workers_per_gpu=0,
    train_dataloader=dict(samples_per_gpu=4),
    train=dict(type='SyntheticPrompts', joint_attention_dim=3584, pooled_projection_dim=None, seq_len=512,
               latent_size=(16, 128, 128), negative=True),
    # train=dict(type='ImagePrompts', cache_dir='data/prompt_cache/qwen', pad_seq_len=512,
    #            negative_prompt_embeds_path='data/prompt_cache/qwen_negative.pt', latent_size=(16, 128, 128)),
)
