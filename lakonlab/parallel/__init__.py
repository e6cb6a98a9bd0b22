from .batch_parallel import gather_latents, init_from_env, parallel_context, shard_batch, shard_bounds  # noqa: F401
